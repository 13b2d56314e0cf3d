#!/usr/bin/env python3
"""Transcribes TestFindTopologyAssignments (pkg/cache/scheduler/tas_cache_test.go:55) into tests/golden/tas_cases.json.
Build-container only (needs /root/reference).  The Go source is parsed and the builder chains are evaluated
symbolically (tools/goparse.py, tools/gointerp.py); nothing of the reference is executed.

Per case: topology levels, nodes (labels, allocatable in Kueue units, taints, readiness), non-TAS pods, flavor node
labels, and per podset the request (TopologyRequest, per-pod requests, count, tolerations, nodeSelector) with the
expected assignment (leaf values + counts) or failure.  Cases outside kb_tas_find's scope are listed under "skipped"."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gointerp  # noqa: E402
from transcribe_tables import eval_function_tables, sym  # noqa: E402
from kueue_b200.api import resource_value  # noqa: E402

PATH = "pkg/cache/scheduler/tas_cache_test.go"
CONSTS = {"tasDataCenterLabel": "cloud.com/datacenter", "tasAIZoneLabel": "cloud.com/aizone", "tasBlockLabel": "cloud.com/topology-block",
          "tasRackLabel": "cloud.com/topology-rack", "tasSubBlockLabel": "cloud.com/topology-subblock"}


def symstr(v):
    """('sym', 'pkg.Name') (tuple, or list after gointerp.strip) -> 'pkg.Name'."""
    if isinstance(v, (list, tuple)) and len(v) == 2 and v[0] == "sym":
        return v[1]
    return v


def k8s(d):
    out = {}
    for k, v in d.items():
        if str(k).startswith("_"):
            continue
        v = symstr(v)
        if isinstance(v, str):
            v = v.split(".")[-1]
            for pre in ("TaintEffect", "TolerationOp"):
                if v.startswith(pre):
                    v = v[len(pre):]
        out[k[0].lower() + k[1:]] = v
    return out


def clean(d):
    return {k: v for k, v in (d or {}).items() if not str(k).startswith("_")}


def norm_node(o):
    n = {"name": o["args"][0], "labels": {}, "allocatable": {}, "taints": [], "ready": False, "unschedulable": False}
    for m, a in o.get("calls", []):
        if m == "Label": n["labels"][a[0]] = a[1]
        elif m == "StatusAllocatable": n["allocatable"] = {r: resource_value(r, q) for r, q in clean(a[0]).items()}
        elif m == "Ready": n["ready"] = True
        elif m == "NotReady": n["ready"] = False
        elif m == "Unschedulable": n["unschedulable"] = True
        elif m == "Taints": n["taints"] = [k8s(t) for t in a]
        elif m == "StatusConditions":
            for c in a:
                if str(symstr(c.get("Type"))).endswith("NodeReady"):
                    n["ready"] = str(symstr(c.get("Status"))).endswith("ConditionTrue")
        else: raise ValueError(f"node builder {m}")
    return n


def norm_pod(o):
    p = {"name": o["args"][0], "node": None, "requests": {}, "phase": "Running"}
    for m, a in o.get("calls", []):
        if m == "NodeName": p["node"] = a[0]
        elif m == "Request": p["requests"][a[0]] = resource_value(a[0], a[1])
        elif m == "StatusPhase": p["phase"] = str(symstr(a[0])).split("Pod")[-1]
        else: raise ValueError(f"pod builder {m}")
    return p


def norm_tr(tr):
    if tr is None:
        return None
    tr = clean(tr)
    if tr.get("PodsetSliceRequiredTopologyConstraints"):
        raise ValueError("multi-layer")
    return {"required": tr.get("Required"), "preferred": tr.get("Preferred"), "unconstrained": bool(tr.get("Unconstrained")),
            "sliceRequiredTopology": tr.get("PodSetSliceRequiredTopology"), "sliceSize": tr.get("PodSetSliceSize")}


def norm_assignment(a):
    if a is None:
        return None
    a = clean(a)
    return {"levels": list(a.get("Levels") or []), "domains": [{"values": list(clean(d).get("Values") or []), "count": clean(d).get("Count")} for d in a.get("Domains") or []]}


def main():
    interp, cases, lines = eval_function_tables(PATH, "TestFindTopologyAssignments", extra_env=dict(CONSTS))
    out = {"source": PATH + ":55", "cases": {}, "skipped": {}}
    for name, ast in cases:
        try:
            tc = interp.ev(ast)
            gates = [str(sym(g)).split(".")[-1] for g in tc.get("enableFeatureGates") or []]
            if any(g in ("TASBalancedPlacement", "TASMultiLayerTopology", "ElasticJobsViaWorkloadSlices", "ElasticJobsViaWorkloadSlicesWithTAS") for g in gates):
                out["skipped"][name] = "feature gate " + ",".join(gates); continue
            pss = []
            for ps in tc.get("podSets") or []:
                ps = clean(ps)
                if ps.get("podSetGroupName") is not None:
                    raise ValueError("leader/worker podset group")
                if ps.get("previousAssignment") is not None:
                    raise ValueError("elastic previousAssignment")
                pss.append({"name": ps.get("podSetName") or "", "topologyRequest": norm_tr(ps.get("topologyRequest")),
                            "requests": {k: int(v) for k, v in clean(ps.get("requests")).items()}, "count": ps.get("count", 0),
                            "tolerations": [k8s(t) for t in ps.get("tolerations") or []],
                            "nodeSelector": clean(ps.get("nodeSelector")) or None,
                            "wantAssignment": norm_assignment(ps.get("wantAssignment")), "wantReason": ps.get("wantReason") or ""})
            case = {"source": f"{PATH}:{lines.get(name, 0)}", "levels": list(tc.get("levels") or []),
                    "nodes": [norm_node(n) for n in tc.get("nodes") or []], "pods": [norm_pod(p) for p in tc.get("pods") or []],
                    "nodeLabels": clean(tc.get("nodeLabels")), "podSets": pss, "gates": gates}
            out["cases"][name] = case
        except Exception as e:  # noqa: BLE001
            out["skipped"][name] = f"{type(e).__name__}: {e}"
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tas_cases.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True, default=str)
    print(f"{len(out['cases'])} cases, {len(out['skipped'])} skipped -> {dst}")
    for k, v in out["skipped"].items():
        print("  skipped:", k[:90], "--", v)


if __name__ == "__main__":
    main()
