#!/usr/bin/env python3
"""Transcribes TestLastSchedulingContext (pkg/scheduler/scheduler_test.go:8569) into tests/golden/last_context_cases.json:
two scheduling cycles with workloads deleted in between; the second cycle starts every pending workload after the
flavors its LastAssignment already tried.  Build-container only; the Go source is parsed, never executed."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gointerp  # noqa: E402
from transcribe_schedule import norm_admission, norm_wl_sched, selector  # noqa: E402
from transcribe_tables import eval_function_tables, norm_cq, sym  # noqa: E402

PATH = "pkg/scheduler/scheduler_test.go"


def main():
    interp, cases, lines = eval_function_tables(PATH, "TestLastSchedulingContext")
    env = interp.env
    lqs = [gointerp.strip(q) for q in env["queues"]]
    out = {"source": PATH + ":8569", "resourceFlavors": [f["name"] for f in env["resourceFlavors"]], "namespaces": {"default": {}}, "cases": {}, "skipped": {}}
    for _, ast in cases:
        tc = interp.ev(ast)
        name = tc["name"]
        try:
            case = {
                "source": f"{PATH}:8569", "enableFairSharing": False, "disablePartialAdmission": False,
                "clusterQueues": [dict(norm_cq(c), namespaceSelector=selector(c), queueingStrategy=str(sym(c.get("queueingStrategy"))).split(".")[-1]) for c in tc["cqs"]],
                "cohorts": [],
                "localQueues": [{"name": q["name"], "ns": q["ns"], "clusterQueue": q["cq"]} for q in lqs],
                "workloads": [norm_wl_sched(w) for w in tc.get("workloads") or []],
                "deleteWorkloads": [f'{d.get("Namespace")}/{d.get("Name")}' for d in tc.get("deleteWorkloads") or []],
                "wantAdmissionsOnSecondSchedule": {k: norm_admission(v) for k, v in (tc.get("wantAdmissionsOnSecondSchedule") or {}).items() if not str(k).startswith("_")},
            }
            out["cases"][name] = case
        except Exception as e:  # noqa: BLE001
            out["skipped"][name] = f"{type(e).__name__}: {e}"
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "last_context_cases.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True, default=str)
    print(f"{len(out['cases'])} cases, {len(out['skipped'])} skipped -> {dst}")
    for k, v in out["skipped"].items():
        print("  skipped:", k, "--", v)


if __name__ == "__main__":
    main()
