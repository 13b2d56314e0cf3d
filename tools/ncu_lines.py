#!/usr/bin/env python3
"""Per-CUDA-source-line totals (warp-stall samples, executed warp instructions) of one kernel from an ncu report:
   python tools/ncu_lines.py gpurun_out/x.ncu-rep [top]        (reads `ncu --page source --print-source cuda,sass`)"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fpath = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name",) or hdr is None: continue
    if r[0].isdigit():  # a source line row with aggregated metrics
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        try: out.append((int(r[si]), int(r[ii]), fpath, int(r[0]), r[1].strip()[:110]))
        except ValueError: pass
tot_s = sum(o[0] for o in out) or 1; tot_i = sum(o[1] for o in out) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}")
print("--- by samples")
for s, i, f, ln, src in sorted(out, key=lambda o: -o[0])[:top]: print(f"{s:6d} {100*s/tot_s:5.1f}% inst {i:8d} {f}:{ln}  {src}")
print("--- by instructions")
for s, i, f, ln, src in sorted(out, key=lambda o: -o[1])[:top]: print(f"{i:8d} {100*i/tot_i:5.1f}% samp {s:6d} {f}:{ln}  {src}")
