#!/usr/bin/env python3
"""Per-kernel device times of one cycle on an arbitrary synthetic snapshot (GPU box):
   python tools/kernel_times.py "synth.make_snapshot(3, W=600, Q=60, heads='one_per_cq', preemption=True, tight=1.1)" [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kueue_b200 import abi, native, synth  # noqa: E402,F401

snap = eval(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
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
