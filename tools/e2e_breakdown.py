"""Wall-clock breakdown of one e2e cycle (GPU box): upload / resident cycle / download / fused kb_run_cycle."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from kueue_b200 import abi, native, synth
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
snap = synth.make_snapshot(cfg, heads="one_per_cq" if cfg >= 3 else "all")
if cfg >= 3:
    snap = synth.compact_to_heads(snap)
ev = native.Evaluator(0)
snap = native.pin_snapshot(snap)
out = native.pin_cycle_out(abi.CycleOut(snap, with_usage=False))
snap.static_generation = 1
def t(f, n=50):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
print("run_cycle ms", t(lambda: ev.run_cycle(snap, out)))
print("upload ms", t(lambda: ev.upload(snap)))
print("resident ms", t(lambda: ev.cycle_resident()))
print("download ms", t(lambda: ev.download(snap, out)))
st = ev.stats(); print("h2d", st.h2d_bytes, st.last_h2d_ms, "d2h", st.d2h_bytes, st.last_d2h_ms, "gpu", st.last_cycle_gpu_ms)
s = snap.as_struct()
print("as_struct ms", t(lambda: snap.as_struct(), 200))
