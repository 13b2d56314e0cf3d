"""Pins the oracle's flavor assignment to the reference's TestAssignFlavors table
(pkg/scheduler/flavorassigner/flavorassigner_test.go:165) with the reference's own stub
preemption oracle (testOracle :145-158).  Fixture: tests/golden/assign_flavors_cases.json
(tools/transcribe_tables.py); workload-slice / TAS cases are out of scope."""
import json
import os

import numpy as np
import pytest

import oracle
from kueue_b200 import abi
from kueue_b200.api import MakePodSet, MakeResourceFlavor, MakeWorkload, flatten
from tests.golden_loader import make_cq, make_cohort

HERE = os.path.dirname(__file__)
DATA = json.load(open(os.path.join(HERE, "golden", "assign_flavors_cases.json")))
CASES = DATA["cases"]


def resource_flavors():
    out = []
    for name, f in DATA["resourceFlavors"].items():
        rf = MakeResourceFlavor(name)
        for k, v in f["nodeLabels"].items():
            rf.NodeLabel(k, v)
        for t in f["taints"]:
            rf.Taint(**t)
        for t in f["tolerations"]:
            rf.Toleration(**t)
        out.append(rf)
    return out


def build(tc):
    cqs = [make_cq(tc["clusterQueue"])]
    if tc["secondaryClusterQueue"]:
        cqs.append(make_cq(tc["secondaryClusterQueue"]))
    pss = []
    for ps in tc["wlPods"]:
        p = MakePodSet(ps["name"], ps["count"])
        for r, q in ps["requests"].items():
            p.Request(r, q)
        p.PodSetGroup(ps.get("group"))
        if ps.get("minCount") is not None:
            p.SetMinimumCount(ps["minCount"])
        for t in ps["tolerations"]:
            p.Toleration(**t)
        if ps["nodeSelector"] is not None:
            p.NodeSelector(ps["nodeSelector"])
        if ps["affinityTerms"]:
            p.RequiredDuringSchedulingIgnoredDuringExecution(ps["affinityTerms"])
        pss.append(p)
    wl = MakeWorkload("wl", "").PodSets(*pss).ClusterQueue(tc["clusterQueue"]["name"]).ReclaimablePods(tc.get("reclaimablePods") or {})
    usage = {tc["clusterQueue"]["name"]: {(f, r): v for f, r, v in tc["clusterQueueUsage"]}}
    if tc["secondaryClusterQueue"]:
        usage[tc["secondaryClusterQueue"]["name"]] = {(f, r): v for f, r, v in tc["secondaryClusterQueueUsage"]}
    flags = abi.FLAGS_DEFAULT | (abi.F_FAIR_SHARING if tc["enableFairSharing"] else 0)
    rfs = resource_flavors()
    extra = [r for _, r, _ in tc["clusterQueueUsage"] + tc["secondaryClusterQueueUsage"] + tc["wantUsage"]] + \
            [r for _, r, _, _ in tc["simulationResult"]]
    names = [f for f, _, _ in tc["clusterQueueUsage"] + tc["secondaryClusterQueueUsage"] + tc["wantUsage"]] + \
            [f for f, _, _, _ in tc["simulationResult"]]
    known = []
    for f in names:
        if f not in known:
            known.append(f)
    snap, idx = flatten(cqs, [], pending=[wl], usage=usage, flags=flags, extra_resources=extra, resource_flavors=rfs,
                        flavors=[f for rg in cqs[0].resource_groups for f in [q.name for q in rg]] + known,
                        reclaimable_pods=tc.get("reclaimablePodsGate", True))
    return snap, idx


@pytest.mark.parametrize("name", list(CASES))
def test_assign_flavors(name):
    tc = CASES[name]
    snap, idx = build(tc)
    sm = np.full(snap.n_fr, -1, np.int8); sb = np.zeros(snap.n_fr, np.int32)
    for f, r, pp, ba in tc["simulationResult"]:
        sm[idx.fr(f, r)] = pp; sb[idx.fr(f, r)] = ba
    got = oracle.assign_stub(snap, 0, sm, sb)
    assert got["mode"] == tc["wantRepMode"], (tc["source"], got)
    for k, ps in enumerate(tc["wantPodSets"]):
        want = {}
        for res, fa in ps["flavors"].items():
            want[res] = (fa["name"], fa["mode"], fa["tried"])
        have = {}
        for r, res in enumerate(idx.resources):
            f = got["flavor"][k, r]
            if f >= 0:
                have[res] = (idx.flavors[f], int(got["res_mode"][k, r]), int(got["tried"][k, r]))
        assert have == want, (tc["source"], k, have, want)
    want_usage = {idx.fr(f, r): v for f, r, v in tc["wantUsage"]}
    have_usage = {fr: int(v) for fr, v in enumerate(got["usage"]) if v >= 0}
    assert have_usage == want_usage, (tc["source"], have_usage, want_usage)
    assert got["borrowing"] == tc["wantBorrowing"], tc["source"]
