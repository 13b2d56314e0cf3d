"""Pins the oracle's whole cycle (nominate -> order -> admit / preempt) to TestSchedule
(pkg/scheduler/scheduler_test.go:69): the quota reservations in the cache after one schedule() call, the workloads
it preempted, and the per-ClusterQueue skipped-preemption counts."""
import json
import os

import pytest

import oracle
from tests.golden_loader import build_schedule_case, schedule_case_result

DOC = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "schedule_cycle_cases.json")))

# The fixture leaves ReclaimWithinCohort empty (no API defaulting in the unit test); the reference compares the
# string with "Never" (flavorassigner.go:1051), so "" enables preemption while borrowing under fair sharing.  The
# flat schema only carries the defaulted policy (webhook default = Never), which cannot express that state.
UNDEFAULTED_POLICY = {"multiple preemptions within cq when fair sharing"}


def schedule_cases():
    return [n for n in DOC["cases"] if n not in UNDEFAULTED_POLICY]


def check_schedule_case(tc, run_cycle):
    snap, idx, entries, admitted = build_schedule_case(DOC, tc)
    out = run_cycle(snap)
    got, preempted, skips = schedule_case_result(snap, idx, entries, admitted, out)
    want = {}
    for key, adm in tc["wantAssignments"].items():
        if key in admitted:
            continue  # reserved before the cycle
        want[key] = {"clusterQueue": adm["clusterQueue"], "podSets": [{"count": p["count"], "flavors": p["flavors"]} for p in adm["podSets"]]}
    assert got == want
    assert preempted == tc["wantPreempted"]
    for cq, n in tc["wantSkippedPreemptions"].items():
        assert skips.get(cq, 0) == n, (cq, skips)
    return snap, out


@pytest.mark.parametrize("name", schedule_cases())
def test_schedule_cycle(name):
    check_schedule_case(DOC["cases"][name], oracle.run_cycle)
