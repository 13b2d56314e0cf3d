#!/usr/bin/env python3
"""Transcribes the reference's Go table tests for the preemption path into JSON fixtures:
   pkg/scheduler/preemption/preemption_test.go               TestPreemption
   pkg/scheduler/preemption/preemption_hierarchical_test.go  TestHierarchicalPreemptions
   pkg/scheduler/preemption/preemption_fair_test.go          TestFairPreemptions
Build-container only (needs /root/reference); the JSON is committed.  The Go sources are
parsed (tools/goparse.py) and the builder chains evaluated symbolically (tools/gointerp.py):
nothing of the reference is executed."""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import goparse  # noqa: E402
import gointerp  # noqa: E402

REF = "/root/reference/"


def function_body(src, name):
    m = re.search(r"\nfunc " + name + r"\(t \*testing\.T\) \{", src)
    start = m.end() - 1
    end = goparse.find_matching(src, start)
    return src[start + 1:end], src[:m.start()].count("\n") + 2


def top_level_statements(body):
    """(name, expr_text, line_offset) for `name := expr` statements at one-tab indent."""
    out = []
    for m in re.finditer(r"\n\t(\w+) :?= ", body):
        out.append((m.group(1), m.end(), body[:m.start()].count("\n") + 1))
    return out


def eval_function_tables(path, func, helpers=None, extra_env=None):
    src = open(REF + path).read()
    body, line0 = function_body(src, func)
    interp = gointerp.Interp(env={"now": gointerp.NOW, **(extra_env or {})}, helpers=helpers or {})
    cases = None
    for name, pos, ln in top_level_statements(body):
        text = body[pos:]
        if name == "cases":
            m = re.match(r"(map\[string\]struct \{|\[\]struct \{)", text)
            if not m:
                continue
            sclose = goparse.find_matching(text, m.end() - 1)
            lit = text[sclose + 1:]
            lit_end = goparse.find_matching(lit, 0)
            ast = goparse.parse_composite_body(lit[:lit_end + 1])
            # keep source lines of each case
            cases = []
            for k, v in ast[2]:
                key = interp.ev(k) if k is not None else None
                cases.append((key, v))
            case_lines = {}
            for mm in re.finditer(r'\n\t\t"((?:[^"\\]|\\.)*)": \{', lit):
                case_lines[mm.group(1)] = line0 + body[:pos].count("\n") + text[:sclose + 1].count("\n") + lit[:mm.start()].count("\n") + 1
            return interp, cases, case_lines
        try:
            p = goparse.Parser(text)
            ast = p.parse_expr()
            interp.env[name] = interp.ev(ast)
        except Exception as e:  # statements we do not need (loops, t.Run, ...)
            pass
    raise RuntimeError("no cases table found in " + func)


POLICY = {"kueue.PreemptionPolicyNever": "Never", "kueue.PreemptionPolicyLowerPriority": "LowerPriority",
          "kueue.PreemptionPolicyLowerOrNewerEqualPriority": "LowerOrNewerEqualPriority", "kueue.PreemptionPolicyAny": "Any",
          "kueue.BorrowWithinCohortPolicyNever": "Never", "kueue.BorrowWithinCohortPolicyLowerPriority": "LowerPriority"}


def sym(v):
    if isinstance(v, tuple) and v and v[0] == "sym":
        return v[1]
    return v


def norm_cq(o):
    p = o.get("preemption") or {}
    bwc = p.get("BorrowWithinCohort") or {}
    ff = o.get("flavorFungibility") or {}
    return {"name": o["name"], "cohort": o.get("cohort"), "resourceGroups": o["resourceGroups"],
            "fairWeight": o.get("fairWeight"),
            "withinClusterQueue": POLICY.get(sym(p.get("WithinClusterQueue")), "Never"),
            "reclaimWithinCohort": POLICY.get(sym(p.get("ReclaimWithinCohort")), "Never"),
            "borrowWithinCohort": POLICY.get(sym(bwc.get("Policy")), "Never") if bwc else "Never",
            "maxPriorityThreshold": bwc.get("MaxPriorityThreshold") if bwc else None,
            "whenCanBorrow": (sym(ff.get("WhenCanBorrow")) or "").split(".")[-1] or None,
            "whenCanPreempt": (sym(ff.get("WhenCanPreempt")) or "").split(".")[-1] or None,
            "preference": (sym(ff.get("Preference")) or "").split(".")[-1] or None if ff.get("Preference") else None}


def norm_cohort(o):
    return {"name": o["name"], "parent": o.get("parent"), "resourceGroups": o["resourceGroups"], "fairWeight": o.get("fairWeight")}


def norm_wl(o):
    conds = []
    for c in o.get("conditions", []):
        conds.append({"type": (sym(c.get("Type")) or "").split(".")[-1].replace("Workload", ""), "status": c.get("Status"),
                      "reason": sym(c.get("Reason"))})
    return {"name": o["name"], "ns": o.get("ns", ""), "priority": o.get("priority", 0), "creation": o.get("creation"),
            "uid": o.get("uid"), "podsets": o["podsets"], "admission": o.get("admission"), "reservedAt": o.get("reservedAt"),
            "conditions": conds}


MODE = {"flavorassigner.NoFit": 0, "flavorassigner.Preempt": 1, "flavorassigner.Fit": 2}


def norm_assignment(a):
    out = []
    for ps in a.get("PodSets", []):
        fl = {}
        for res, fa in (ps.get("Flavors") or {}).items():
            if str(res).startswith("_"):
                continue
            fl[res] = {"name": fa.get("Name"), "mode": MODE.get(sym(fa.get("Mode")), 0)}
        out.append({"name": ps.get("Name", "main"), "count": ps.get("Count", 1), "flavors": fl})
    return out


def single_podset_assignment(interp, args):
    return {"PodSets": [{"Name": "main", "Flavors": args[0], "Count": 1}], "_type": "flavorassigner.Assignment"}


REASONS = {"kueue.InClusterQueueReason": "InClusterQueue", "kueue.InCohortReclamationReason": "InCohortReclamation",
           "kueue.InCohortFairSharingReason": "InCohortFairSharing",
           "kueue.InCohortReclaimWhileBorrowingReason": "InCohortReclaimWhileBorrowing"}


def target_key_reason(interp, args):
    return {"_tkr": [args[0], REASONS.get(sym(args[1]), sym(args[1]))]}


def preemption_cases(path, func, fair=False):
    interp, cases, lines = eval_function_tables(path, func, helpers={"singlePodSetAssignment": single_podset_assignment,
                                                                       "targetKeyReason": target_key_reason})
    out = {}
    for key, v in cases:
        tc = interp.ev(v)
        want = {}
        for w in tc.get("wantWorkloads", []) or []:
            for c in w.get("conditions", []):
                if (gointerp.strip(c).get("Type") or ("", ""))[1:] and sym(c.get("Type")) == "kueue.WorkloadPreempted":
                    want[w["name"]] = sym(c.get("Reason"))
        if isinstance(tc.get("wantPreempted"), list):  # sets.New(targetKeyReason("/name", reason)...) style
            for item in tc["wantPreempted"]:
                wkey, reason = item["_tkr"]
                want[wkey.split("/")[-1]] = reason
        assignment = norm_assignment(tc.get("assignment") or {})
        if fair:  # preemption_fair_test.go:995-1007: cpu on assignmentFlavor (default "default"), mode Preempt
            assignment = [{"name": "main", "count": 1, "flavors": {"cpu": {"name": tc.get("assignmentFlavor") or "default", "mode": 1}}}]
        case = {
            "source": f"{path}:{lines.get(key, 0)}",
            "clusterQueues": [norm_cq(c) for c in (tc.get("clusterQueues") or [])],
            "cohorts": [norm_cohort(c) for c in (tc.get("cohorts") or [])],
            "admitted": [norm_wl(w) for w in (tc.get("admitted") or [])],
            "incoming": norm_wl(tc["incoming"]) if tc.get("incoming") else None,
            "targetCQ": tc.get("targetCQ"),
            "assignment": assignment,
            "strategies": [sym(x).split(".")[-1] for x in (tc.get("strategies") or [])],
            "want": want,
            "wantPreempted": tc.get("wantPreempted") if isinstance(tc.get("wantPreempted"), int) else None,
        }
        out[key] = gointerp.strip(case)
    return out


if __name__ == "__main__":
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    c = preemption_cases("pkg/scheduler/preemption/preemption_test.go", "TestPreemption")
    json.dump(c, open(os.path.join(dst, "preemption_cases.json"), "w"), indent=1, default=str)
    print("TestPreemption:", len(c), "cases; ignored builder methods:", sorted(gointerp.IGNORED))
    c = preemption_cases("pkg/scheduler/preemption/preemption_hierarchical_test.go", "TestHierarchicalPreemptions")
    json.dump(c, open(os.path.join(dst, "preemption_hierarchical_cases.json"), "w"), indent=1, default=str)
    print("TestHierarchicalPreemptions:", len(c), "cases; ignored:", sorted(gointerp.IGNORED))
    c = preemption_cases("pkg/scheduler/preemption/preemption_fair_test.go", "TestFairPreemptions", fair=True)
    json.dump(c, open(os.path.join(dst, "preemption_fair_cases.json"), "w"), indent=1, default=str)
    print("TestFairPreemptions:", len(c), "cases; ignored:", sorted(gointerp.IGNORED))


# ---------------------------------------------------------------------------
# TestAssignFlavors  pkg/scheduler/flavorassigner/flavorassigner_test.go:165
# ---------------------------------------------------------------------------
PP = {"preemptioncommon.NoCandidates": 1, "preemptioncommon.Preempt": 2, "preemptioncommon.Reclaim": 3}
AMODE = {"NoFit": 0, "Preempt": 1, "Fit": 2}


def _frq(d):
    out = []
    for k, v in (d or {}).items():
        if str(k).startswith("_"):
            continue
        kk = dict(k)
        out.append([kk.get("Flavor"), kk.get("Resource"), v])
    return out


def norm_podset(o):
    return {"name": o["name"], "count": o["count"], "minCount": o.get("minCount"), "group": o.get("group"), "requests": o["requests"],
            "tolerations": o.get("tolerations") or [], "nodeSelector": o.get("nodeSelector"), "affinityTerms": o.get("affinityTerms")}


def assign_cases():
    path = "pkg/scheduler/flavorassigner/flavorassigner_test.go"
    interp, cases, lines = eval_function_tables(path, "TestAssignFlavors", extra_env={"Fit": ("sym", "Fit"), "Preempt": ("sym", "Preempt"), "NoFit": ("sym", "NoFit")})
    flavors = {k: gointerp.strip(v) for k, v in interp.env["resourceFlavors"].items() if not str(k).startswith("_")}
    out, skipped = {}, {}
    src = open(REF + path).read()
    lit_text = {}
    ks = sorted((ln, k) for k, ln in lines.items())
    src_lines = src.split("\n")
    for i, (ln, k) in enumerate(ks):
        end = ks[i + 1][0] if i + 1 < len(ks) else ln + 400
        lit_text[k] = "\n".join(src_lines[ln - 1:end - 1])
    for key, v in cases:
        tc = interp.ev(v)
        if tc.get("elasticJobsViaWorkloadSlicesEnabled") or tc.get("preemptWorkloadSlice"):
            skipped[key] = "workload slices (alpha gate ElasticJobsViaWorkloadSlices, out of scope)"
            continue
        gates = {str(sym(k)).split(".")[-1]: bool(v) for k, v in (tc.get("featureGates") or {}).items() if not str(k).startswith("_")}
        if set(gates) - {"ReclaimablePods"}:
            skipped[key] = "non-default feature gates"
            continue
        wa = tc.get("wantAssignment") or {}
        want_ps = []
        tas = False
        for ps in wa.get("PodSets", []) or []:
            fl = {}
            for res, fa in (ps.get("Flavors") or {}).items():
                if str(res).startswith("_"):
                    continue
                fl[res] = {"name": fa.get("Name"), "mode": AMODE.get(sym(fa.get("Mode")), 0), "tried": fa.get("TriedFlavorIdx", 0)}
            if ps.get("TopologyAssignment") or ps.get("DelayedTopologyRequest"):
                tas = True
            want_ps.append({"name": ps.get("Name"), "count": ps.get("Count"), "flavors": fl})
        if tas:
            skipped[key] = "TAS"
            continue
        sim = []
        for k, r in (tc.get("simulationResult") or {}).items():
            if str(k).startswith("_"):
                continue
            kk = dict(k)
            if isinstance(r, list):
                pp, ba = r[0], (r[1] if len(r) > 1 else 0)
            else:
                pp, ba = r.get("preemptionPossiblity"), r.get("borrowingAfterSimulation", 0)
            sim.append([kk.get("Flavor"), kk.get("Resource"), PP[sym(pp)], ba])
        cq = tc["clusterQueue"]
        case = {
            "source": f"{path}:{lines.get(key, 0)}",
            "wlPods": [norm_podset(p) for p in tc.get("wlPods") or []],
            "clusterQueue": norm_cq(cq),
            "clusterQueueUsage": _frq(tc.get("clusterQueueUsage")),
            "secondaryClusterQueue": norm_cq(tc["secondaryClusterQueue"]) if tc.get("secondaryClusterQueue") else None,
            "secondaryClusterQueueUsage": _frq(tc.get("secondaryClusterQueueUsage")),
            "enableFairSharing": bool(tc.get("enableFairSharing")),
            "reclaimablePods": {p["Name"]: p["Count"] for p in gointerp.strip(tc.get("wlReclaimablePods") or [])},
            "reclaimablePodsGate": gates.get("ReclaimablePods", True),
            "simulationResult": sim,
            "wantRepMode": AMODE[sym(tc.get("wantRepMode"))] if tc.get("wantRepMode") is not None else 0,
            "wantPodSets": want_ps,
            "wantUsage": _frq(((wa.get("Usage") or {}).get("Quota")) or {}),
            "wantBorrowing": wa.get("Borrowing", 0),
        }
        out[key] = gointerp.strip(case)
    return out, skipped, flavors


if __name__ == "__main__":
    c, skipped, flavors = assign_cases()
    json.dump({"resourceFlavors": flavors, "cases": c, "skipped": skipped}, open(os.path.join(dst, "assign_flavors_cases.json"), "w"), indent=1, default=str)
    print("TestAssignFlavors:", len(c), "cases transcribed,", len(skipped), "skipped; ignored:", sorted(gointerp.IGNORED))
