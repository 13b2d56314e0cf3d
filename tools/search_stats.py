#!/usr/bin/env python3
"""Per-kernel device times and target-search counters of one cycle (GPU box): python tools/search_stats.py [config] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kueue_b200 import abi, native, synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
snap = synth.make_snapshot(cfg, heads="one_per_cq")
snap = synth.compact_to_heads(snap)
ev = native.Evaluator(0)
ev.upload(snap)
ev.set_profile(True)
for _ in range(reps):
    ev.cycle_resident()
st = ev.stats()
print("cycle ms", st.last_cycle_gpu_ms, "launches", st.kernel_launches)
for i, nm in enumerate(abi.KERNEL_NAMES):
    if st.kernel_ms[i] > 0:
        print(f"  {nm:26s} {st.kernel_ms[i]:9.3f} ms")
ss = list(st.search_stat)
print("searches", ss[0], "records", ss[1], "visited", ss[2], "removed", ss[3], "records of multi-column searches", ss[4])
if ss[0]:
    print("per search: records %.0f, cycles load %.0f classify %.0f greedy %.0f" % (ss[1] / ss[0], ss[5] / ss[0], ss[6] / ss[0], ss[7] / ss[0]))
