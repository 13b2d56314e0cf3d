"""Drain mode (kueue_b200/drain.py): iterated cycles over whole queues.  CPU: the loop's bookkeeping on the oracle;
GPU: the same drain through the device library is identical cycle by cycle."""
import numpy as np
import pytest

import oracle
from kueue_b200 import abi, synth
from kueue_b200.drain import drain

CASES = [
    lambda: synth.make_snapshot(1),                                                   # 100 workloads, 10 cohort-less ClusterQueues
    lambda: synth.make_snapshot(2, W=3000, Q=30, podsets_max=2),                      # StrictFIFO: a blocked head stops its queue
    lambda: synth.make_snapshot(3, W=2000, Q=40),                                     # BestEffortFIFO + fair sharing in flat cohorts
    lambda: synth.make_snapshot(3, W=1500, Q=60, partial=True, podsets_max=2),                # partial admission: reduced counts enter the usage
]


def _check(snap, res):
    a = snap.arrays
    Q, FR = snap.n_cq, snap.n_fr
    assert res.cycles >= 1 and len(res.heads) == res.cycles
    adm = np.concatenate(res.admitted) if res.admitted else np.zeros(0, np.int64)
    assert len(np.unique(adm)) == len(adm), "a workload was admitted twice"
    for heads, dec in zip(res.heads, res.decisions):
        cqs = a["wl_cq"][heads]
        assert len(np.unique(cqs)) == len(cqs), "more than one head of a ClusterQueue in a cycle"
    # every head is the best not-yet-admitted, not-set-aside workload of its queue: priorities never increase along a queue
    seen = {}
    for heads in res.heads:
        for w in heads:
            c = int(a["wl_cq"][w])
            key = (-int(a["wl_priority"][w]), int(a["wl_ts"][w]), int(a["wl_uid"][w]))
            if c in seen and seen[c][1] != int(w):
                assert seen[c][0] < key, "queue order violated"
            seen[c] = (key, int(w))
    assert (res.cq_usage >= a["cq_usage"].reshape(Q, FR)).all()
    assert len(res.admitted[-1]) == 0 or res.cycles == 10_000 or all(len(h) for h in res.heads)


@pytest.mark.parametrize("make", CASES)
def test_drain_on_oracle(make):
    snap = make()
    res = drain(snap, oracle.run_cycle)
    _check(snap, res)
    assert res.n_admitted > 0 and res.cycles > 1
    again = drain(snap, oracle.run_cycle)
    assert again.cycles == res.cycles and all(np.array_equal(x, y) for x, y in zip(again.decisions, res.decisions))


# TestBestEffortFIFORequeueIfNotPresent (pkg/cache/queue/cluster_queue_test.go:749-815), transcribed by hand: requeue
# reason + LastAssignment.LastTriedFlavorIdx -> does the workload land in the inadmissible set of a BestEffortFIFO queue?
_REQUEUE_CASES = {
    "failure after nomination": ("FailedAfterNomination", None, False),
    "namespace doesn't match": ("NamespaceMismatch", None, True),
    "didn't fit and no pending flavors": ("", [{"memory": -1}, {"cpu": -1, "memory": -1}], True),
    "didn't fit but pending flavors": ("", [{"cpu": -1, "memory": 0}, {"memory": 1}], False),
}


@pytest.mark.parametrize("name", list(_REQUEUE_CASES))
def test_requeue_rule_matches_reference_table(name):
    from kueue_b200.drain import requeue_goes_inadmissible
    reason, last, want = _REQUEUE_CASES[name]
    rows = None
    if last is not None:  # LastTriedFlavorIdx as [podset][resource] rows, absent resources = -1 like the flat tables
        res = ["cpu", "memory"]
        rows = np.array([[ps.get(r, -1) for r in res] for ps in last], np.int8)
    assert requeue_goes_inadmissible(False, reason, rows) == want
    # StrictFIFO requeues immediately unless the namespace does not match (cluster_queue.go:622-624): the head stays
    assert requeue_goes_inadmissible(True, reason, rows) == (reason == "NamespaceMismatch" and not (rows is not None and (rows != -1).any()))


def test_first_drain_cycle_is_the_reference_cycle():
    snap = synth.make_snapshot(3, W=2000, Q=40, heads="one_per_cq")
    one = oracle.run_cycle(synth.compact_to_heads(snap))
    res = drain(synth.make_snapshot(3, W=2000, Q=40), oracle.run_cycle, max_cycles=1)
    # same heads (queues.Heads) in ClusterQueue order, same decisions
    assert np.array_equal(np.sort(res.heads[0]), np.sort(snap.arrays["heads"]))
    by_cq = {int(snap.arrays["wl_cq"][w]): int(d) for w, d in zip(snap.arrays["heads"], one.decision)}
    assert [by_cq[int(snap.arrays["wl_cq"][w])] for w in res.heads[0]] == res.decisions[0].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("make", CASES)
def test_drain_device_equals_oracle(make):
    from kueue_b200 import native
    ev = native.Evaluator(0)
    try:
        snap = make()
        want = drain(snap, oracle.run_cycle)
        got = drain(snap, ev.run_cycle)
        assert got.cycles == want.cycles
        for k in range(want.cycles):
            assert np.array_equal(got.heads[k], want.heads[k]) and np.array_equal(got.decisions[k], want.decisions[k]), k
        assert np.array_equal(got.cq_usage, want.cq_usage)
    finally:
        ev.close()


DRAIN_DEVICE_CASES = CASES + [
    lambda: synth.make_snapshot(2, W=3000, Q=30, preemption=True, tight=1.2),                  # within-ClusterQueue preemption: admitted tables grow and are re-ranked on the device
    lambda: synth.make_snapshot(4, W=2000, Q=100, tight=1.06),                                  # hierarchical cohorts, reclaim
    lambda: _with_hashes(synth.make_snapshot(3, W=4000, Q=40)),                                 # scheduling-hash bulk move (BestEffortFIFO)
]


def _with_hashes(snap):
    rng = np.random.default_rng(3)
    snap.set("wl_sched_hash", rng.integers(0, 6, snap.n_wl))  # few classes, 0 = unknown
    return snap


@pytest.mark.gpu
@pytest.mark.parametrize("make", DRAIN_DEVICE_CASES)
def test_kb_run_drain_equals_host_definition(make):
    """kb_run_drain (queue layer on the device) == kueue_b200/drain.py iterating the oracle, cycle by cycle."""
    from kueue_b200 import native
    ev = native.Evaluator(0)
    try:
        snap = make()
        cap = 40 * (snap.n_adm + snap.n_wl) + 10000
        want = drain(snap, lambda s: oracle.run_cycle(s, cap))
        got = ev.run_drain(snap)
        assert got.n_cycles == want.cycles
        cyc = got.cycles()
        for k in range(want.cycles):
            assert np.array_equal(cyc[k][0], want.heads[k]), k
            assert np.array_equal(cyc[k][1], want.decisions[k]), k
        assert np.array_equal(got.cq_usage, want.cq_usage)
        adm = np.concatenate(want.admitted) if want.admitted else np.zeros(0, np.int64)
        assert got.n_admitted == len(adm) and np.array_equal(np.flatnonzero(got.wl_admit_cycle >= 0), np.sort(adm))
    finally:
        ev.close()


@pytest.mark.gpu
def test_kb_run_drain_full_size_config3():
    """cfg3 at full size: 1M pending workloads reach the device; the drain equals the iterated oracle."""
    from kueue_b200 import native
    ev = native.Evaluator(0)
    try:
        snap = synth.make_snapshot(3)
        want = drain(snap, oracle.run_cycle, max_cycles=6)
        got = ev.run_drain(snap, max_cycles=6)
        assert got.n_cycles == want.cycles
        cyc = got.cycles()
        for k in range(want.cycles):
            assert np.array_equal(cyc[k][0], want.heads[k]) and np.array_equal(cyc[k][1], want.decisions[k]), k
        assert np.array_equal(got.cq_usage, want.cq_usage)
    finally:
        ev.close()
