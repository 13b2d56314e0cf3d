"""Pins the host-side queue-order timestamp, the oracle's classical iterator order and resourcesToReserve to the
reference's own table tests (tests/golden/schedule_cases.py)."""
import numpy as np
import pytest

import oracle
from kueue_b200 import abi
from kueue_b200.api import MakeClusterQueue, MakeFlavorQuotas, MakeWorkload, flatten, queue_order_timestamp
from tests.golden.schedule_cases import ENTRY_ORDERING_CASES, HOUR, NOW, QUEUE_ORDER_TS_CASES, RESERVE_CASES
from tests.schedule_golden import entry_ordering_snapshot, order_of


@pytest.mark.parametrize("name", list(QUEUE_ORDER_TS_CASES))
def test_queue_order_timestamp(name):  # workload_test.go:698
    cond, want = QUEUE_ORDER_TS_CASES[name]
    w = MakeWorkload("name", "ns").Creation(NOW)
    if cond:
        w.Condition(cond[0], cond[1], cond[2], NOW + HOUR)
    for ordering, which in want.items():
        assert queue_order_timestamp(w, ordering) == (NOW if which == "creation" else NOW + HOUR), ordering


@pytest.mark.parametrize("name", list(ENTRY_ORDERING_CASES))
def test_entry_ordering(name):  # scheduler_test.go:8291
    tc = ENTRY_ORDERING_CASES[name]
    snap, idx, want_borrow = entry_ordering_snapshot(tc)
    out = oracle.run_cycle(snap)
    assert (out.decision == abi.DEC_ASSUMED).all()
    assert (out.borrow == want_borrow).all(), "fixture must reproduce the Borrowing levels of the reference entries"
    assert order_of(out, tc) == tc["want"]


@pytest.mark.parametrize("name", list(RESERVE_CASES))
def test_resources_to_reserve(name):  # scheduler_test.go:9705
    tc = RESERVE_CASES[name]
    cq = (MakeClusterQueue("cq").Cohort("eng")
          .ResourceGroup(MakeFlavorQuotas("on-demand").Resource("memory", "100"), MakeFlavorQuotas("spot").Resource("memory", "0", "100"))
          .ResourceGroup(MakeFlavorQuotas("model-a").Resource("gpu", "10", "0"), MakeFlavorQuotas("model-b").Resource("gpu", "10", "5")))
    snap, idx = flatten([cq], usage={"cq": dict(tc["cq_usage"])})
    usage = np.full(snap.n_fr, -1, np.int64)
    for (f, r), v in tc["usage"].items():
        usage[idx.fr(f, r)] = v
    got = oracle.resources_to_reserve(snap, 0, abi.MODE_PREEMPT if tc["mode"] == "Preempt" else abi.MODE_FIT, tc["borrowing"], usage)
    want = np.full(snap.n_fr, -1, np.int64)
    for (f, r), v in tc["want"].items():
        want[idx.fr(f, r)] = v
    assert got.tolist() == want.tolist()


def test_entry_comparer_less():  # scheduler_test.go:9577
    from tests.golden.schedule_cases import ENTRY_LESS_CASES, S
    for name, tc in ENTRY_LESS_CASES.items():
        def op(e):
            _, created, borrowing, drs = e
            ratio, weight = drs if drs is not None else (0.0, 0.0)
            return (borrowing, 0, NOW + created * S, ratio, weight)
        assert oracle.entry_less(abi.FLAGS_DEFAULT, op(tc["a"]), op(tc["b"])) == tc["want"], name


def _two_cq_snapshot(admitted, flags=abi.FLAGS_DEFAULT):
    from kueue_b200.api import MakeAdmission
    cqs = [MakeClusterQueue(n).Cohort("co").ResourceGroup(MakeFlavorQuotas("default").Resource("cpu", "100")) for n in ("preemptor", "other")]
    return flatten(cqs, admitted=admitted, flags=flags, now_ns=NOW)


def test_satisfies_preemption_policy():  # preemption_policy_test.go:32
    from kueue_b200.api import MakeAdmission
    from tests.golden.schedule_cases import POLICY_CASES
    pol = {"Never": abi.POLICY_NEVER, "LowerPriority": abi.POLICY_LOWER_PRIORITY,
           "LowerOrNewerEqualPriority": abi.POLICY_LOWER_OR_NEWER_EQUAL_PRIORITY, "Any": abi.POLICY_ANY}
    MIN = 60 * 10**9
    for name, ((pp, pt), (cp, ct), policy, buffer_gate, want) in POLICY_CASES.items():
        cand = MakeWorkload("candidate").Priority(cp).Creation(NOW + ct * MIN).ReserveQuota(MakeAdmission("other").Assignment("cpu", "default", "1"), NOW)
        flags = abi.FLAGS_DEFAULT | (abi.F_TS_PREEMPTION_BUFFER if buffer_gate else 0)
        snap, idx = _two_cq_snapshot([cand], flags)
        assert oracle.satisfies_policy(snap, pp, NOW + pt * MIN, 0, pol[policy]) == want, name


def test_candidates_ordering():  # preemption_test.go:4525
    from kueue_b200.api import MakeAdmission
    from tests.golden.schedule_cases import CANDIDATE_ORDER_CASES, S
    for name, tc in CANDIDATE_ORDER_CASES.items():
        adm = []
        for uid, (wname, cq, prio, reserved, evicted) in enumerate(tc["cands"]):
            w = MakeWorkload(wname).Priority(prio).UID(uid + 1).ReserveQuota(MakeAdmission(cq).Assignment("cpu", "default", "1"),
                                                                         None if reserved is None else NOW + reserved * S)
            if evicted:
                w.Evicted()
            adm.append(w)
        snap, idx = _two_cq_snapshot(adm)
        got = [tc["cands"][i][0] for i in oracle.sort_candidates(snap, idx.cqs.index("preemptor"))]
        assert got == tc["want"], name


def test_podset_reducer_search():  # podset_reducer_test.go:26
    from tests.golden.schedule_cases import REDUCER_CASES
    from tests.schedule_golden import reducer_snapshot
    for name, tc in REDUCER_CASES.items():
        snap, idx = reducer_snapshot(tc)
        out = oracle.run_cycle(snap)
        if tc["found"]:
            assert out.decision[0] == abi.DEC_ASSUMED and int(out.ps_count.sum()) == tc["count"], name
            assert all(int(c) >= (m if m is not None else n) for c, (n, m) in zip(out.ps_count, tc["podsets"])), name
        else:
            assert out.decision[0] != abi.DEC_ASSUMED, name


def test_usage_with_lending_limit():  # snapshot_test.go:1131 — cohort usage is a function of the ClusterQueue usages
    from tests.golden.schedule_cases import LENDING_CASES
    from tests.schedule_golden import lending_snapshot
    for name, (remaining, (cohort, a, b)) in LENDING_CASES.items():
        snap, idx = lending_snapshot(remaining)
        out = oracle.tree_eval(snap)
        fr = idx.fr("default", "cpu")
        assert (int(out.usage[idx.node("lend"), fr]), int(out.usage[idx.node("lend-a"), fr]), int(out.usage[idx.node("lend-b"), fr])) == (cohort, a, b), name
        assert int(out.subtree_quota[idx.node("lend"), fr]) == 10_000, name


def test_is_preferred():  # flavorassigner_test.go:3885 (granular modes: 2 preempt, 4 fit)
    FIT, PREEMPT = 4, 2
    cases = {
        "feature gate disabled prioritises preemption": ((FIT, 0), (PREEMPT, 0), abi.PREF_UNSET, True),
        "explicit BorrowingOverPreemption prioritises borrowing distance": ((PREEMPT, 1), (FIT, 2), abi.PREF_BORROWING_OVER_PREEMPTION, False),
        "explicit PreemptionOverBorrowing prioritises lower preemption": ((PREEMPT, 1), (FIT, 2), abi.PREF_PREEMPTION_OVER_BORROWING, True),
        "explicit PreemptionOverBorrowing breaks borrowing ties with preemption": ((PREEMPT, 1), (FIT, 1), abi.PREF_PREEMPTION_OVER_BORROWING, False),
    }
    for name, (a, b, pref, want) in cases.items():
        assert oracle.is_preferred(a, b, pref) == want, name


def _last_assignment_snapshot(cq_generation, wl_generation):
    """A workload that tried flavor index 0 last time: it resumes at index 1 only while the ClusterQueue's
    AllocatableResourceGeneration has not moved past the generation it remembered (lastAssignmentOutdated,
    flavorassigner.go:749-752; TestLastAssignmentOutdated flavorassigner_test.go:3705)."""
    from kueue_b200.api import MakePodSet
    cq = (MakeClusterQueue("cq").Generation(cq_generation).FlavorFungibility("TryNextFlavor", "TryNextFlavor")
          .ResourceGroup(MakeFlavorQuotas("one").Resource("cpu", "10"), MakeFlavorQuotas("two").Resource("cpu", "10")))
    w = MakeWorkload("wl").ClusterQueue("cq").PodSets(MakePodSet("main", 1).Request("cpu", "1")).LastAssignment([{"cpu": 0}], wl_generation)
    return flatten([cq], pending=[w], now_ns=NOW)


def test_last_assignment_outdated():
    for name, cq_gen, wl_gen, flavor in (("Cluster queue allocatableResourceIncreasedGen increased", 1, 0, "one"),
                                         ("AllocatableResourceGeneration not increased", 0, 0, "two")):
        snap, idx = _last_assignment_snapshot(cq_gen, wl_gen)
        out = oracle.run_cycle(snap)
        assert out.decision[0] == abi.DEC_ASSUMED, name
        assert idx.flavors[int(out.ps_flavor[0, idx.resources.index("cpu")])] == flavor, name


def test_strict_fifo_head_selection():  # cluster_queue_test.go:892 (host-side queues.Heads mirror)
    from kueue_b200.api import select_heads
    from tests.golden.schedule_cases import STRICT_FIFO_CASES, S
    for name, (a, b, ordering, want) in STRICT_FIFO_CASES.items():
        ws = []
        for wname, (created, prio, evicted_at) in (("w1", a), ("w2", b)):
            w = MakeWorkload(wname).ClusterQueue("cq").Priority(prio).Creation(NOW + created * S)
            if evicted_at is not None:
                w.Condition("Evicted", True, "PodsReadyTimeout", NOW + evicted_at * S)
            ws.append(w)
        for order in (ws, ws[::-1]):  # push order must not matter
            assert select_heads(order, ordering)[0].name == want, name
