"""TestLastSchedulingContext (pkg/scheduler/scheduler_test.go:8569) as two cycles through any single-cycle evaluator:
cycle 1, apply its decisions the way schedule() does (admissions into the cache, Evicted condition on preemption targets,
LastAssignment of the entries that stay pending), delete the listed workloads, cycle 2, compare the admissions in the cache."""
import copy
import json
import os

import numpy as np

from kueue_b200 import abi
from tests.golden_loader import BASE, build_schedule_case, schedule_case_result

DOC = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "last_context_cases.json")))


def run_two_cycles(tc, run_cycle):
    specs = {f'{w["ns"]}/{w["name"]}': copy.deepcopy(w) for w in tc["workloads"]}
    last = {}
    outs = []
    for cycle in (0, 1):
        tcx = dict(tc, workloads=list(specs.values()), wantAssignments={}, wantPreempted=[], wantSkippedPreemptions={})
        snap, idx, entries, admitted = build_schedule_case(DOC, tcx)
        R = snap.n_resource
        for e, key in enumerate(entries):  # LastAssignment kept by the queue between cycles (workload.go:107-110,178-191)
            if key in last:
                wl = int(snap.arrays["heads"][e])
                snap.arrays["wl_last_gen"][wl] = 0  # ClusterQueueGeneration of the attempt; no spec change in between
                r0, r1 = int(snap.arrays["wl_ps_start"][wl]), int(snap.arrays["wl_ps_start"][wl + 1])
                snap.arrays["ps_last_tried"].reshape(-1, R)[r0:r1] = last[key]
        out = run_cycle(snap)
        outs.append((snap, out))
        got, preempted, _ = schedule_case_result(snap, idx, entries, admitted, out)
        for e, key in enumerate(entries):
            wl = int(snap.arrays["heads"][e])
            r0, r1 = int(snap.arrays["wl_ps_start"][wl]), int(snap.arrays["wl_ps_start"][wl + 1])
            if out.decision[e] == abi.DEC_ASSUMED:
                sp = specs[key]
                a = got[key]
                sp["admission"] = {"cq": a["clusterQueue"], "podsets": [
                    {"name": ps["name"], "count": pa["count"], "assignments": {r: [f, ps["requests"][r], pa["count"]] for r, f in pa["flavors"].items()}}
                    for ps, pa in zip(sp["podsets"], a["podSets"])]}
                sp["reservedAt"] = BASE
                last.pop(key, None)
            elif out.decision[e] == abi.DEC_PREEMPTING:
                last.pop(key, None)  # scheduler.go:345: the next attempt tries all the flavors
            else:
                last[key] = np.asarray(out.ps_tried_idx).reshape(-1, R)[r0:r1].copy()  # scheduler.go:494
        for key in preempted:  # IssuePreemptions -> Evicted condition; the workload keeps its quota until it is deleted
            specs[key].setdefault("conditions", []).append({"type": "Evicted", "status": "True", "reason": "Preempted"})
        if cycle == 0:
            for key in tc["deleteWorkloads"]:
                specs.pop(key)
                last.pop(key, None)
    admissions = {}
    for key, sp in specs.items():
        if sp.get("admission"):
            admissions[key] = {"clusterQueue": sp["admission"]["cq"],
                               "podSets": [{"name": p["name"], "count": p.get("count", 1), "flavors": {r: v[0] for r, v in p["assignments"].items()}} for p in sp["admission"]["podsets"]]}
    return admissions, outs


def check(tc, run_cycle):
    got, outs = run_two_cycles(tc, run_cycle)
    want = {k: {"clusterQueue": v["clusterQueue"], "podSets": [{"name": p["name"], "count": p["count"], "flavors": p["flavors"]} for p in v["podSets"]]}
            for k, v in tc["wantAdmissionsOnSecondSchedule"].items()}
    assert got == want, (got, want)
    return outs
