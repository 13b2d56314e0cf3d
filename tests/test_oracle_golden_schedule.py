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
