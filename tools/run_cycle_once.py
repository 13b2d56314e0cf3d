"""Runs a few cycles of a synthetic snapshot (for ncu captures of a single kernel):
   python tools/run_cycle_once.py <config> <W> <Q> [cycles]"""
import sys
sys.path.insert(0, ".")
from kueue_b200 import abi, native, synth
cfg, W, Q = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 3
snap = synth.make_snapshot(cfg, W=W, Q=Q, heads="one_per_cq" if cfg >= 3 else "all")
if cfg >= 3:
    snap = synth.compact_to_heads(snap)
ev = native.Evaluator(0)
out = abi.CycleOut(snap, 40 * snap.n_adm + 10000, with_usage=False)
for _ in range(n):
    ev.run_cycle(snap, out)
st = ev.stats()
print("heads", snap.n_heads, "adm", snap.n_adm, "gpu_ms", st.last_cycle_gpu_ms, "targets", out.n_targets)
