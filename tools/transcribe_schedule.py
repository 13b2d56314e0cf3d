#!/usr/bin/env python3
"""Transcribes TestSchedule (pkg/scheduler/scheduler_test.go:69) into tests/golden/schedule_cycle_cases.json.
Build-container only (needs /root/reference).  The Go source is parsed and the builder chains are evaluated
symbolically (tools/goparse.py, tools/gointerp.py); nothing of the reference is executed.

Per case: feature gates, ClusterQueues (base + additional), Cohorts, LocalQueues, workloads, and what one
scheduling cycle must produce: wantAssignments (quota reservations in the cache after the cycle), the workloads
the cycle preempted (Evicted/Preempted conditions in wantWorkloads) and wantSkippedPreemptions."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gointerp  # noqa: E402
from transcribe_tables import eval_function_tables, norm_cq, norm_cohort, norm_wl, sym  # noqa: E402

PATH = "pkg/scheduler/scheduler_test.go"


def selector(o):
    sel = o.get("namespaceSelector", "unset")
    if sel == "unset":
        return "everything"  # MakeClusterQueue default: &metav1.LabelSelector{} (wrappers.go)
    if sel is None:
        return "nothing"
    exprs = []
    for e in sel.get("MatchExpressions") or []:
        exprs.append({"key": e.get("Key"), "operator": str(sym(e.get("Operator"))).split("LabelSelectorOp")[-1], "values": list(e.get("Values") or [])})
    labels = {k: v for k, v in (sel.get("MatchLabels") or {}).items() if not str(k).startswith("_")}
    return {"matchExpressions": exprs, "matchLabels": labels}


def norm_wl_sched(o):
    w = norm_wl(o)
    w["queue"] = o.get("queue")
    for c, src in zip(w["conditions"], o.get("conditions", [])):
        c["lastTransitionTime"] = src.get("LastTransitionTime") if isinstance(src.get("LastTransitionTime"), int) else None
    return w


def norm_admission(a):
    if isinstance(a, gointerp.Obj) or "podsets" in a:  # MakeAdmission(...)
        return {"clusterQueue": a["cq"], "podSets": [{"name": p["name"], "count": p.get("count", 1),
                                                       "flavors": {r: v[0] for r, v in p["assignments"].items()}} for p in a["podsets"]]}
    out = {"clusterQueue": a.get("ClusterQueue"), "podSets": []}
    for p in a.get("PodSetAssignments") or []:
        if "assignments" in p:  # MakePodSetAssignment(...)
            out["podSets"].append({"name": p["name"], "count": p.get("count", 1), "flavors": {r: v[0] for r, v in p["assignments"].items()}})
        else:  # kueue.PodSetAssignment{Name:, Flavors:, Count:}
            fl = {r: f for r, f in (p.get("Flavors") or {}).items() if not str(r).startswith("_")}
            out["podSets"].append({"name": p.get("Name", "main"), "count": p.get("Count", 1), "flavors": fl})
    return out


def main():
    interp, cases, lines = eval_function_tables(PATH, "TestSchedule")
    env = interp.env
    base_cqs = list(env["clusterQueues"])
    base_lqs = [gointerp.strip(q) for q in env["queues"]]
    flavors = [f["name"] for f in env["resourceFlavors"]]
    out = {"source": PATH + ":69", "resourceFlavors": flavors,
           "namespaces": {"default": {}, "eng-alpha": {"dep": "eng"}, "eng-beta": {"dep": "eng"}, "eng-gamma": {"dep": "eng"},
                          "sales": {"dep": "sales"}, "lend": {"dep": "lend"}},  # scheduler_test.go (client objects of the runner)
           "cases": {}, "skipped": {}}
    for name, ast in cases:
        try:
            tc = interp.ev(ast)
        except Exception as e:  # noqa: BLE001
            out["skipped"][name] = f"not evaluable: {type(e).__name__}: {e}"
            continue
        if tc.get("enableElasticJobsViaWorkloadSlice"):
            out["skipped"][name] = "workload slices (out of scope)"; continue
        if tc.get("objects"):
            out["skipped"][name] = "needs LimitRange / extra API objects (validation before the cycle)"; continue
        if any(ps.get("limits") for w in tc.get("workloads") or [] for ps in w["podsets"]):
            out["skipped"][name] = "request/limit validation before the cycle"; continue
        if tc.get("admissionError") is not None:
            out["skipped"][name] = "injects an API error after the cycle's decision"; continue
        cqs = base_cqs + list(tc.get("additionalClusterQueues") or [])
        lqs = base_lqs + [gointerp.strip(q) for q in tc.get("additionalLocalQueues") or []]
        wls = [norm_wl_sched(w) for w in tc.get("workloads") or []]
        want_wls = [norm_wl_sched(w) for w in tc.get("wantWorkloads") or []]
        evicted_before = {f'{w["ns"]}/{w["name"]}' for w in wls if any(c["type"] == "Evicted" and c["status"] == "True" for c in w["conditions"])}
        preempted = sorted(f'{w["ns"]}/{w["name"]}' for w in want_wls
                           if any(c["type"] == "Evicted" and c["status"] == "True" and str(c["reason"]).endswith("Preempted") or
                                  str(c["reason"]) == "kueue.WorkloadEvictedByPreemption" for c in w["conditions"])
                           and f'{w["ns"]}/{w["name"]}' not in evicted_before)
        wa = tc.get("wantAssignments") or {}
        case = {
            "source": f"{PATH}:{lines.get(name, 0)}",
            "enableFairSharing": bool(tc.get("enableFairSharing")), "disablePartialAdmission": bool(tc.get("disablePartialAdmission")),
            "clusterQueues": [dict(norm_cq(c), namespaceSelector=selector(c), queueingStrategy=str(sym(c.get("queueingStrategy"))).split(".")[-1]) for c in cqs],
            "cohorts": [norm_cohort(c) for c in tc.get("cohorts") or []],
            "localQueues": [{"name": q["name"], "ns": q["ns"], "clusterQueue": q["cq"]} for q in lqs],
            "workloads": wls,
            "wantAssignments": {k: norm_admission(v) for k, v in wa.items() if not str(k).startswith("_")},
            "wantPreempted": preempted,
            "wantSkippedPreemptions": {k: v for k, v in (tc.get("wantSkippedPreemptions") or {}).items() if not str(k).startswith("_")},
        }
        out["cases"][name] = case
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "schedule_cycle_cases.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True, default=str)
    print(f"{len(out['cases'])} cases, {len(out['skipped'])} skipped -> {dst}")
    for k, v in out["skipped"].items():
        print("  skipped:", k, "--", v)
    print("ignored builder methods:", sorted(gointerp.IGNORED))


if __name__ == "__main__":
    main()
