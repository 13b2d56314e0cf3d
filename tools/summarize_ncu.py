"""Summarise an ncu report (read here, no GPU) into profiles/:
  python tools/summarize_ncu.py gpurun_out/r01_full_cfg3.ncu-rep cfg3 r01
writes profiles/<tag>_ncu_<cfg>.txt (per-kernel table + top stall sites) and updates profiles/traffic.json
(dram__bytes_read.sum + dram__bytes_write.sum per launch, used by bench.py's roofline.traffic)."""
import csv, io, json, os, subprocess, sys

rep, cfg, tag = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%")]
idx = [(hdr.index(m) if m in hdr else -1, n) for m, n in want]


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


lines = [f"# ncu --set full --clock-control none, {os.path.basename(rep)} ({cfg}); one steady-state cycle, cold-cache serialised launches",
         "# " + " | ".join(f"{n}[{units[i] if i >= 0 else ''}]" for i, n in idx)]
traffic = {}
for r in rows[2:]:
    vals = [(r[i] if i >= 0 else "") for i, _ in idx]
    lines.append(" | ".join(v[:48] for v in vals))
    name = vals[0].split("(")[0].replace("void ", "").split("<")[0]
    ird, iwr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    traffic.setdefault(name, []).append(to_bytes(r[ird], units[ird]) + to_bytes(r[iwr], units[iwr]))

# top stall sites of the longest kernel
longest = max(rows[2:], key=lambda r: float(r[hdr.index("gpu__time_duration.sum")]))[hdr.index("Kernel Name")]
kname = longest.split("(")[0].replace("void ", "").split("<")[0]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kname}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    sh = srows[1]
    si, so = sh.index("# Samples"), sh.index("Source")
    stall = [i for i, h in enumerate(sh) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in srows[2:2 + (len(srows) - 2) // 1] if len(r) > si and r[si].isdigit()]
    half = body[:len(body) // 2] if len(body) > 2 and body[0][so] == body[len(body) // 2][so] else body  # the CSV repeats the listing
    tot = sum(int(r[si]) for r in half) or 1
    lines.append(f"\n# warp-stall samples of {kname} (SASS, top 15 of {tot} samples)")
    agg = {}
    for r in half:
        for i in stall:
            agg[sh[i]] = agg.get(sh[i], 0) + int(r[i] or 0)
    lines.append("# by reason: " + ", ".join(f"{k}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
    for r in sorted(half, key=lambda r: -int(r[si]))[:15]:
        top = max(stall, key=lambda i: int(r[i] or 0))
        lines.append(f"{int(r[si]):6d}  {r[so].strip()[:60]:60s}  {sh[top]}")
out = os.path.join(ROOT, "profiles", f"{tag}_ncu_{cfg}.txt")
open(out, "w").write("\n".join(lines) + "\n")
tj = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(tj)) if os.path.exists(tj) else {}
t[cfg] = {k: sum(v) / len(v) for k, v in traffic.items()}
t[cfg]["k_nominate"] = t[cfg].get("k_nominate_coop", t[cfg].get("k_nominate"))
if "k_cycle_flat" not in t[cfg] and "k_cycle_root" in t[cfg]:
    t[cfg]["k_cycle_flat"] = t[cfg]["k_cycle_root"]  # kb_stats slot of the fused per-root cycle: whichever of the two kernels ran
json.dump(t, open(tj, "w"), indent=1, sort_keys=True)
print(open(out).read())
