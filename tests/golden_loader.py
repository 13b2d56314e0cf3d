"""Builds flat snapshots from the JSON fixtures produced by tools/transcribe_tables.py."""
from kueue_b200 import abi
from kueue_b200.api import (MakeAdmission, MakeClusterQueue, MakeCohort, MakeFlavorQuotas, MakePodSet, MakeWorkload,
                            flatten, resource_value)

BASE = 1_700_000_000_000_000_000


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


def make_cq(spec):
    cq = MakeClusterQueue(spec["name"])
    if spec.get("cohort"):
        cq.Cohort(spec["cohort"])
    if spec.get("fairWeight") is not None:
        cq.FairWeight(float(resource_value("x", spec["fairWeight"])) if False else _weight(spec["fairWeight"]))
    cq.Preemption(withinClusterQueue=spec.get("withinClusterQueue", "Never"), reclaimWithinCohort=spec.get("reclaimWithinCohort", "Never"),
                  borrowWithinCohort=spec.get("borrowWithinCohort", "Never"), maxPriorityThreshold=spec.get("maxPriorityThreshold"))
    wb, wp = spec.get("whenCanBorrow"), spec.get("whenCanPreempt")
    if wb or wp or spec.get("preference"):
        cq.FlavorFungibility(wb or "MayStopSearch", wp or "TryNextFlavor", spec.get("preference"))
    return _quotas(cq, spec["resourceGroups"])


def _weight(w):
    from fractions import Fraction
    s = str(w)
    if s.endswith("m"):
        return float(Fraction(s[:-1]) / 1000)
    return float(Fraction(s))


def make_cohort(spec):
    co = MakeCohort(spec["name"])
    if spec.get("parent"):
        co.Parent(spec["parent"])
    if spec.get("fairWeight") is not None:
        co.FairWeight(_weight(spec["fairWeight"]))
    return _quotas(co, spec["resourceGroups"])


def make_workload(spec, uid):
    w = MakeWorkload(spec["name"], spec.get("ns", "")).Priority(spec.get("priority", 0)).UID(uid)
    w.Creation(spec.get("creation") or BASE)
    pss = []
    for ps in spec["podsets"]:
        p = MakePodSet(ps["name"], ps["count"])
        for r, q in ps["requests"].items():
            p.Request(r, q)
        if ps.get("minCount") is not None:
            p.SetMinimumCount(ps["minCount"])
        pss.append(p)
    w.PodSets(*pss)
    adm = spec.get("admission")
    if adm:
        a = MakeAdmission(adm["cq"])
        first = True
        for psa in adm["podsets"]:
            if not first:
                a.PodSet()
            first = False
            for r, v in psa["assignments"].items():
                qty = resource_value(r, v[1]) * (v[2] if len(v) > 2 else 1)
                # MakeAdmission stores quantities through resource_value again: pass int64 units
                a.podsets[-1][r] = (v[0], _Raw(qty))
        w.ReserveQuota(a, spec.get("reservedAt"))
    for c in spec.get("conditions", []):
        if c.get("type") == "Evicted" and c.get("status") == "True":
            w.Evicted()
    return w


class _Raw(int):
    """An int64 already in Kueue units (bypasses quantity parsing)."""
    _raw_units = True


def build_preemption_case(tc, flags=abi.FLAGS_DEFAULT):
    cqs = [make_cq(c) for c in tc["clusterQueues"]]
    cohorts = [make_cohort(c) for c in tc["cohorts"]]
    # uid order = name order (the reference compares Obj.UID strings; tests leave them empty or name-like)
    names = sorted(w["name"] for w in tc["admitted"])
    admitted = [make_workload(w, 1000 + names.index(w["name"])) for w in tc["admitted"]]
    inc = make_workload(tc["incoming"], 1).ClusterQueue(tc["targetCQ"])
    extra = {r for ps in tc["assignment"] for r in ps["flavors"]}
    flv = [f["name"] for ps in tc["assignment"] for f in ps["flavors"].values() if f["name"]]
    snap, idx = flatten(cqs, cohorts, pending=[inc], admitted=admitted, flags=flags, now_ns=BASE, extra_resources=extra,
                        flavors=None)
    return snap, idx
