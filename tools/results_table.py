#!/usr/bin/env python3
"""Markdown table of the committed bench lines (profiles/<tag>_bench_cfg*.json): python tools/results_table.py [tag]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def last_json(path):
    line = None
    for ln in open(path):
        ln = ln.strip()
        if ln.startswith("{"):
            line = ln
    return json.loads(line) if line else None


ref = None
p = os.path.join(ROOT, "profiles", f"{tag}_bench_reference_cfg3.json")
if os.path.exists(p):
    ref = last_json(p)
print("| config | decisions / cycle | device ms / cycle | decisions/s (device) | decisions/s (e2e, host buffers) | dominant kernel: ms, fraction of measured HBM peak | CPU port, 1 thread (decisions/s) | e2e ÷ CPU |")
print("|---|---|---|---|---|---|---|---|")
for cfg in (2, 3, 4, 5):
    p = os.path.join(ROOT, "profiles", f"{tag}_bench_cfg{cfg}.json")
    if not os.path.exists(p):
        continue
    d = last_json(p)
    if not d:
        continue
    r, c = d["roofline"], d.get("cpu_baseline") or {}
    cpu = c.get("value")
    ratio = f"{d['e2e']['value'] / cpu:,.0f}×" if cpu else "—"
    print(f"| cfg{cfg} | {d['config']['decisions_per_step_per_gpu']:,} | {d['ms_per_step']:.4f} | {d['value']:.3g} | {d['e2e']['value']:.3g} | "
          f"`{r['kernel']}` {r['kernel_ms']:.4f} ms, {r['frac']:.3f} | {cpu:,.0f} | {ratio} |" if cpu else
          f"| cfg{cfg} | {d['config']['decisions_per_step_per_gpu']:,} | {d['ms_per_step']:.4f} | {d['value']:.3g} | {d['e2e']['value']:.3g} | `{r['kernel']}` {r['kernel_ms']:.4f} ms, {r['frac']:.3f} | — | — |")
    if d.get("e2e_incremental"):
        i = d["e2e_incremental"]
        print(f"| cfg{cfg}, usage deltas | | | | {i['value']:.3g} ({i['h2d_bytes_per_step'] / 1e6:.2f} MB H2D instead of {d['e2e']['h2d_bytes_per_step'] / 1e6:.2f}) | | | |")
    if d.get("drain"):
        dr = d["drain"]
        cpu_d = (dr.get("cpu") or {}).get("decisions_per_s")
        print(f"| cfg{cfg} drain ({dr['pending']:,} pending, {dr['cycles']} cycles, {dr['admitted']:,} admitted) | {dr['decisions']:,} total | {dr['device_ms']:.2f} total | {dr['decisions_per_s_device']:.3g} | {dr['decisions_per_s_e2e']:.3g} | | "
              + (f"{cpu_d:,.0f} | {dr['decisions_per_s_e2e'] / cpu_d:,.0f}× |" if cpu_d else "— | — |"))
if ref:
    print(f"\nReference arm (`bench.py --impl reference`, cfg3): {ref['value']:,.0f} decisions/s, {ref['cpu_baseline']['kind']}, {ref['cpu_baseline']['cores']} thread.")
