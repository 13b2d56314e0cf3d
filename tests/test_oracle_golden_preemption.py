"""Pins the oracle's classical / hierarchical preemption (GetTargets) to the reference's
TestPreemption table (pkg/scheduler/preemption/preemption_test.go:64, 36 cases) — fixture
tests/golden/preemption_cases.json produced by tools/transcribe_tables.py."""
import json
import os

import numpy as np
import pytest

import oracle
from kueue_b200 import abi
from tests.golden_loader import build_preemption_case

HERE = os.path.dirname(__file__)
REASON = {abi.REASON_IN_CLUSTER_QUEUE: "InClusterQueue", abi.REASON_IN_COHORT_RECLAMATION: "InCohortReclamation",
          abi.REASON_IN_COHORT_FAIR_SHARING: "InCohortFairSharing",
          abi.REASON_IN_COHORT_RECLAIM_WHILE_BORROWING: "InCohortReclaimWhileBorrowing"}


def assignment_arrays(tc, snap, idx):
    R = snap.n_resource
    np_ = len(tc["assignment"])
    fl = np.full((np_, R), -1, np.int8); md = np.full((np_, R), -1, np.int8); cnt = np.zeros(np_, np.int32)
    for k, ps in enumerate(tc["assignment"]):
        cnt[k] = ps["count"]
        for res, fa in ps["flavors"].items():
            r = idx.resources.index(res)
            fl[k, r] = idx.flavors.index(fa["name"]); md[k, r] = fa["mode"]
    return fl, md, cnt


def run_case(tc, flags=abi.FLAGS_DEFAULT):
    snap, idx = build_preemption_case(tc, flags)
    fl, md, cnt = assignment_arrays(tc, snap, idx)
    got = oracle.get_targets(snap, 0, fl, md, cnt)
    return {idx.admitted[a]: REASON[r] for a, r in got}


CASES = json.load(open(os.path.join(HERE, "golden", "preemption_cases.json")))


def check_targets(got, tc):
    """`want` lists the workloads that gained a Preempted condition (wantWorkloads); targets that were
    already evicted get no new condition (preemption.go:183-187) and only count in wantPreempted."""
    n = tc["wantPreempted"] or 0
    assert len(got) == n, (tc["source"], got)
    for name, reason in tc["want"].items():
        assert got.get(name) == reason, (tc["source"], got)
    if len(tc["want"]) == n:
        assert got == tc["want"], tc["source"]


@pytest.mark.parametrize("name", list(CASES))
def test_preemption(name):
    tc = CASES[name]
    check_targets(run_case(tc), tc)
