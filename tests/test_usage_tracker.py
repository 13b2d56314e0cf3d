"""Host side of the incremental snapshot (kueue_b200.api.UsageTracker): which form of the usage table a cycle carries."""
import numpy as np

from kueue_b200 import abi, synth
from kueue_b200.api import UsageTracker


def _snap(gen, usage=None):
    s = synth.make_snapshot(3, W=300, Q=30, heads="one_per_cq")
    s.static_generation = gen
    if usage is not None:
        s.set("cq_usage", usage)
    return s


def test_first_call_sends_the_full_table_and_asks_to_keep_it():
    t = UsageTracker()
    s = t.prepare(_snap(7))
    assert s.flags & abi.F_USAGE_RESIDENT and "usage_delta_cq" not in s.arrays


def test_touched_and_changed_rows_become_the_delta():
    t = UsageTracker()
    first = _snap(7)
    base = np.array(first.arrays["cq_usage"]).reshape(first.n_cq, first.n_fr).copy()
    t.prepare(first)
    nxt = base.copy(); nxt[4] += 5; nxt[9] -= 1          # row 9 changes without a touch: found by the diff
    t.touch(4); t.touch(11)                              # row 11 touched but unchanged: still sent (harmless)
    s = t.prepare(_snap(7, nxt))
    assert not (s.flags & abi.F_USAGE_RESIDENT)
    assert s.arrays["usage_delta_cq"].tolist() == [4, 9, 11]
    assert np.array_equal(s.arrays["usage_delta_rows"].reshape(3, -1), nxt[[4, 9, 11]])
    s = t.prepare(_snap(7, nxt))                         # nothing changed since: an empty delta
    assert len(s.arrays["usage_delta_cq"]) == 0 and s.as_struct().n_usage_delta == 0


def test_new_static_generation_or_reset_starts_over():
    t = UsageTracker()
    t.prepare(_snap(7))
    assert t.prepare(_snap(8)).flags & abi.F_USAGE_RESIDENT            # ClusterQueue / Cohort specs changed
    assert not (t.prepare(_snap(8)).flags & abi.F_USAGE_RESIDENT)
    t.reset()                                                          # library error -> resident table unknown
    assert t.prepare(_snap(8)).flags & abi.F_USAGE_RESIDENT
    for _ in range(2):                                                 # no static generation: nothing can stay resident
        s = t.prepare(_snap(0))
        assert not (s.flags & abi.F_USAGE_RESIDENT) and "usage_delta_cq" not in s.arrays
