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


# ---------------------------------------------------------------------------
# TestSchedule (scheduler_test.go:69) cases from tools/transcribe_schedule.py
# ---------------------------------------------------------------------------
def _selector_matches(sel, labels):
    if sel == "everything":
        return True
    if sel == "nothing":
        return False
    for k, v in sel.get("matchLabels", {}).items():
        if labels.get(k) != v:
            return False
    for e in sel.get("matchExpressions", []):
        has, val = e["key"] in labels, labels.get(e["key"])
        op = e["operator"]
        if op == "In" and not (has and val in e["values"]): return False
        if op == "NotIn" and has and val in e["values"]: return False
        if op == "Exists" and not has: return False
        if op == "DoesNotExist" and has: return False
    return True


def build_schedule_case(doc, tc, capture=None):
    """One scheduling cycle of a TestSchedule case.  Returns (snap, idx, keys of the entries, info) where the host
    side does what the queue manager / cache do before schedule(): LocalQueue -> ClusterQueue resolution, inactive
    ClusterQueues (missing ResourceFlavor, cache/clusterqueue.go), namespace selector (scheduler.go:589-595), one head
    per ClusterQueue by (priority desc, queue-order timestamp asc) (queue/cluster_queue.go:667-680)."""
    from kueue_b200.api import queue_order_timestamp
    flags = abi.FLAGS_DEFAULT
    if tc["enableFairSharing"]:
        flags |= abi.F_FAIR_SHARING
    if tc["disablePartialAdmission"]:
        flags &= ~abi.F_PARTIAL_ADMISSION
    known = set(doc["resourceFlavors"])
    active = [c for c in tc["clusterQueues"] if all(f["flavor"] in known for rg in c["resourceGroups"] for f in rg)]
    cq_by_name = {c["name"]: c for c in active}
    cqs = [make_cq(c) for c in active]
    cohorts = [make_cohort(c) for c in tc["cohorts"]]
    lq = {(q["ns"], q["name"]): q["clusterQueue"] for q in tc["localQueues"]}
    names = sorted(f'{w["ns"]}/{w["name"]}' for w in tc["workloads"])
    admitted, pending_by_cq = [], {}
    for spec in tc["workloads"]:
        key = f'{spec["ns"]}/{spec["name"]}'
        w = make_workload(spec, 1000 + names.index(key))
        w.name = key
        for c in spec.get("conditions", []):
            if c.get("lastTransitionTime") is not None:
                w.Condition(c["type"], c["status"] == "True", str(c.get("reason")).split(".")[-1].replace("WorkloadEvictedBy", ""), c["lastTransitionTime"])
        if spec.get("admission") and spec.get("reservedAt") is not None:
            if spec["admission"]["cq"] in cq_by_name:
                admitted.append(w)
            continue
        cq = lq.get((spec["ns"], spec.get("queue")))
        if cq is None or cq not in cq_by_name:
            continue  # unknown LocalQueue / ClusterQueue, or inactive ClusterQueue: never reaches the cycle
        pending_by_cq.setdefault(cq, []).append((w, spec))
    pending, dropped = [], []
    from kueue_b200.api import select_heads
    for cq, lst in pending_by_cq.items():
        w = select_heads([t[0].ClusterQueue(cq) for t in lst])[0]  # uid = rank of the key, the reference's UID tie-break
        spec = next(sp for ww, sp in lst if ww is w)
        if not _selector_matches(cq_by_name[cq]["namespaceSelector"], doc["namespaces"].get(spec["ns"], {})):
            dropped.append(w.name)  # nominate: "Workload namespace doesn't match ClusterQueue selector" -> never an entry
            continue
        pending.append(w.ClusterQueue(cq))
    snap, idx = flatten(cqs, cohorts, pending=pending, admitted=admitted, flags=flags, now_ns=BASE, flavors=list(doc["resourceFlavors"]))
    if capture is not None:
        capture.update(cqs=cqs, cohorts=cohorts, pending=pending, admitted=admitted, flags=flags)
    return snap, idx, [w.name for w in pending], [w.name for w in admitted]


def schedule_case_result(snap, idx, entry_keys, admitted_keys, out):
    """(admissions of the cycle {key: {cq, podsets: [{name?, count, flavors}]}}, preempted keys, skips per CQ)."""
    import numpy as np
    R = snap.n_resource
    a = snap.arrays
    res = {}
    for e, key in enumerate(entry_keys):
        if out.decision[e] != abi.DEC_ASSUMED:
            continue
        wl = int(a["heads"][e])
        pss = []
        for row in range(int(a["wl_ps_start"][wl]), int(a["wl_ps_start"][wl + 1])):
            fl = {idx.resources[r]: idx.flavors[int(out.ps_flavor[row, r])] for r in range(R) if out.ps_flavor[row, r] >= 0}
            pss.append({"count": int(out.ps_count[row]), "flavors": fl})
        res[key] = {"clusterQueue": idx.cqs[int(a["wl_cq"][wl])], "podSets": pss}
    pre = set()
    skips = {}
    for e in range(len(entry_keys)):
        wl = int(a["heads"][e])
        if out.decision[e] == abi.DEC_PREEMPTING:
            pre.update(admitted_keys[int(out.tgt_adm[k])] for k in range(int(out.tgt_start[e]), int(out.tgt_start[e + 1])))
        # skippedPreemptions scheduler.go:321-332: overlapping targets, or a Preempt-mode entry that no longer fits
        if out.decision[e] == abi.DEC_SKIPPED_OVERLAP or (out.decision[e] == abi.DEC_SKIPPED_NO_FIT and out.mode[e] == abi.MODE_PREEMPT):
            cq = idx.cqs[int(a["wl_cq"][wl])]
            skips[cq] = skips.get(cq, 0) + 1
    return res, sorted(pre), skips
