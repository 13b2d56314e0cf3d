#!/usr/bin/env python3
"""TestReclaimBeforePriorityPreemption (flavorassigner_test.go:3383) and TestHierarchical (:3755) case tables ->
tests/golden/assign_extra_cases.json.  The fixed environment of both tests (ClusterQueues / Cohorts built in the test
body) is restated by hand in tests/test_oracle_golden_assign_extra.py."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gointerp  # noqa: E402
from transcribe_tables import AMODE, PP, _frq, eval_function_tables, norm_podset, sym  # noqa: E402

PATH = "pkg/scheduler/flavorassigner/flavorassigner_test.go"
out = {}
for func in ("TestReclaimBeforePriorityPreemption", "TestHierarchical"):
    interp, cases, lines = eval_function_tables(PATH, func, extra_env={"Fit": ("sym", "Fit"), "Preempt": ("sym", "Preempt"), "NoFit": ("sym", "NoFit")})
    tab = {}
    for key, v in cases:
        tc = interp.ev(v)
        sim = []
        for k, r in (tc.get("simulationResult") or {}).items():
            if str(k).startswith("_"):
                continue
            kk = dict(k)
            sim.append([kk.get("Flavor"), kk.get("Resource"), PP[sym(r[0])], r[1] if len(r) > 1 else 0])
        ff = tc.get("flavorFungibility") or None
        tab[key] = gointerp.strip({
            "source": f"{PATH}:{lines.get(key, 0)}",
            "podSet": norm_podset(tc["workloadRequests"]),
            "testClusterQueueUsage": _frq(tc.get("testClusterQueueUsage")),
            "otherClusterQueueUsage": _frq(tc.get("otherClusterQueueUsage")),
            "flavorFungibility": None if not ff else {k: str(sym(x)).split(".")[-1] for k, x in ff.items() if not str(k).startswith("_")},
            "simulationResult": sim,
            "wantMode": AMODE[sym(tc["wantMode"])],
            "wantAssignment": {k: x for k, x in (tc.get("wantAssigment") or {}).items() if not str(k).startswith("_")},
        })
    out[func] = tab
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "assign_extra_cases.json")
json.dump(out, open(dst, "w"), indent=1, sort_keys=True, default=str)
print({k: len(v) for k, v in out.items()}, "->", dst)
