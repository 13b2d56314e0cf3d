"""Pins the oracle's DominantResourceShare to the reference's
TestDominantResourceShare (pkg/cache/scheduler/fair_sharing_test.go:37, 15 cases;
fixture produced by tools/transcribe_drs.py)."""
import json
import os

import numpy as np
import pytest

import oracle
from kueue_b200.api import MakeClusterQueue, MakeCohort, MakeFlavorQuotas, flatten

CASES = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "drs_cases.json")))


def _quotas(holder, rgs):
    for rg in rgs:
        fqs = []
        for f in rg:
            fq = MakeFlavorQuotas(f["flavor"])
            for r in f["resources"]:
                fq.Resource(r["name"], r["nominal"], r["borrowingLimit"] or "", r["lendingLimit"] or "")
            fqs.append(fq)
        holder.ResourceGroup(*fqs)
    return holder


def build(tc):
    cqs = []
    for spec in (tc["clusterQueue"], tc["lendingClusterQueue"]):
        if spec is None:
            continue
        cq = MakeClusterQueue(spec["name"])
        if spec["cohort"]:
            cq.Cohort(spec["cohort"])
        if spec["fairWeight"] is not None:
            cq.FairWeight(float(spec["fairWeight"]))
        cqs.append(_quotas(cq, spec["resourceGroups"]))
    cohorts = []
    for spec in tc["cohorts"]:
        co = MakeCohort(spec["name"])
        if spec["parent"]:
            co.Parent(spec["parent"])
        if spec["fairWeight"] is not None:
            co.FairWeight(float(spec["fairWeight"]))
        cohorts.append(_quotas(co, spec["resourceGroups"]))
    usage = {"cq": {(f, r): v for f, r, v in tc["usage"]}}
    extra = [r for _, r, _ in tc["usage"] + tc["flvResQ"]]
    flv = [f for f, _, _ in tc["usage"] + tc["flvResQ"]]
    return cqs, cohorts, usage, extra, flv


@pytest.mark.parametrize("name", list(CASES))
def test_dominant_resource_share(name):
    tc = CASES[name]
    cqs, cohorts, usage, extra, flv = build(tc)
    snap, idx = flatten(cqs, cohorts, usage=usage, extra_resources=extra, flavors=None)
    wl_req = None
    if tc["flvResQ"]:
        wl_req = np.zeros(snap.n_fr, np.int64)
        for f, r, v in tc["flvResQ"]:
            wl_req[idx.fr(f, r)] = v
    out = oracle.tree_eval(snap, wl_req)
    for w in tc["want"]:
        n = idx.node(w["name"])
        assert int(out.drs_rounded[n]) == w["drValue"], (w, int(out.drs_rounded[n]))
        got_name = idx.resources[out.drs_resource[n]] if out.drs_resource[n] >= 0 else ""
        assert got_name == w["drName"], w
        assert bool(out.drs_borrowing[n]) == w["borrowing"], w
