"""Pins flavor assignment in cohorts to TestReclaimBeforePriorityPreemption (flavorassigner_test.go:3383) and
TestHierarchical (:3755).  Case tables: tests/golden/assign_extra_cases.json (tools/transcribe_assign_extra.py); the
fixed ClusterQueues / Cohorts of each test body (:3505-3528, :3793-3827) are restated here.  Both tests use the
reference's stub preemption oracle (testOracle, :145-158)."""
import json
import os

import numpy as np
import pytest

import oracle
from kueue_b200 import abi
from kueue_b200.api import MakeClusterQueue, MakeCohort, MakeFlavorQuotas, MakePodSet, MakeWorkload, flatten

DATA = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "assign_extra_cases.json")))


def _fung(cq, ff):
    if ff:
        cq.FlavorFungibility(ff.get("WhenCanBorrow", "MayStopSearch"), ff.get("WhenCanPreempt", "TryNextFlavor"))
    return cq


def env_reclaim(ff):
    rg = lambda q: [MakeFlavorQuotas(f).Resource("compute", q).Resource("gpu", q) for f in ("uno", "due", "tre")]  # noqa: E731
    test = (MakeClusterQueue("test-clusterqueue").Cohort("cohort").Preemption("LowerPriority", "LowerPriority")
            .FlavorFungibility("MayStopSearch", "TryNextFlavor").ResourceGroup(*rg("10")))
    other = MakeClusterQueue("other-clusterqueue").Cohort("cohort").ResourceGroup(*rg("0"))
    return [_fung(test, ff), other], []


def env_hierarchical(ff):
    cohorts = [MakeCohort("three").ResourceGroup(MakeFlavorQuotas("three").Resource("cpu", "4")),
               MakeCohort("two").Parent("three").ResourceGroup(MakeFlavorQuotas("two").Resource("cpu", "4")),
               MakeCohort("one").Parent("two").ResourceGroup(MakeFlavorQuotas("one").Resource("cpu", "4"))]
    rg = lambda: [MakeFlavorQuotas(f).Resource("cpu", "0") for f in ("one", "two", "three")]  # noqa: E731
    test = (MakeClusterQueue("test-clusterqueue").Cohort("one").Preemption("LowerPriority", "LowerPriority")
            .ResourceGroup(*rg()).FlavorFungibility("MayStopSearch", "TryNextFlavor"))
    other = MakeClusterQueue("other-clusterqueue").Cohort("two").ResourceGroup(*rg())
    return [_fung(test, ff), other], cohorts


ENVS = {"TestReclaimBeforePriorityPreemption": env_reclaim, "TestHierarchical": env_hierarchical}


def build_assign_extra(func, tc):
    cqs, cohorts = ENVS[func](tc["flavorFungibility"])
    ps = MakePodSet(tc["podSet"]["name"], tc["podSet"]["count"])
    for r, q in tc["podSet"]["requests"].items():
        ps.Request(r, q)
    wl = MakeWorkload("wl", "").PodSets(ps).ClusterQueue("test-clusterqueue")
    usage = {"test-clusterqueue": {(f, r): v for f, r, v in tc["testClusterQueueUsage"]},
             "other-clusterqueue": {(f, r): v for f, r, v in tc["otherClusterQueueUsage"]}}
    return flatten(cqs, cohorts, pending=[wl], usage=usage)


def _cases():
    return [pytest.param(func, name, id=f"{func[4:]}:{name[:50]}") for func, tab in DATA.items() for name in tab]


@pytest.mark.parametrize("func,name", _cases())
def test_assign_in_cohorts(func, name):
    tc = DATA[func][name]
    snap, idx = build_assign_extra(func, tc)
    sm = np.full(snap.n_fr, -1, np.int8); sb = np.zeros(snap.n_fr, np.int32)
    for f, r, pp, ba in tc["simulationResult"]:
        sm[idx.fr(f, r)] = pp; sb[idx.fr(f, r)] = ba
    got = oracle.assign_stub(snap, 0, sm, sb)
    assert got["mode"] == tc["wantMode"], tc["source"]
    have = {idx.resources[r]: idx.flavors[got["flavor"][0, r]] for r in range(snap.n_resource) if got["flavor"][0, r] >= 0}
    assert have == tc["wantAssignment"], tc["source"]
