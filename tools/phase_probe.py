import sys; sys.path.insert(0,'.')
from kueue_b200 import abi, native, synth
snap = synth.compact_to_heads(synth.make_snapshot(3, heads="one_per_cq"))
ev = native.Evaluator(0); ev.upload(snap); ev.set_profile(True)
for _ in range(3): ev.cycle_resident()
st = ev.stats(); print("cycle ms", st.last_cycle_gpu_ms, [ (abi.KERNEL_NAMES[i], round(st.kernel_ms[i],4)) for i in range(16) if st.kernel_ms[i]>0])
print("phase cycles (k_cycle_flat: stage, entries+tree, gather+available, fair+nominate, keys+expand, rank+thresholds, ordered loop, update+publish):", list(st.search_stat)[:8])
