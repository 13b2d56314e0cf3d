"""The C++ host mirror (include/kueue_b200_host.hpp) against the Python one on the TestSchedule scenarios:
identical kb_snapshot tables (CPU), identical decisions to the oracle through Scheduler::schedule (GPU)."""
import json
import os
import subprocess

import numpy as np
import pytest

from kueue_b200 import abi
from tests import cpp_host_gen
from tests.golden_loader import BASE, build_schedule_case
from tests.test_oracle_golden_schedule_cycle import DOC, schedule_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORK = os.path.join(ROOT, "tests", "cpp", "_build")


def _scenarios():
    out = []
    for name in schedule_cases():
        parts = {}
        snap, idx, entries, admitted = build_schedule_case(DOC, DOC["cases"][name], capture=parts)
        out.append((name, snap, idx, parts))
    return out


@pytest.fixture(scope="module")
def driver():
    if not os.path.exists(os.path.join(ROOT, "kueue_b200", "libkueue_b200.so")):
        import __graft_entry__
        __graft_entry__.build()
    sc = _scenarios()
    cpp = [cpp_host_gen.scenario_cpp(k, p["cqs"], p["cohorts"], p["pending"], p["admitted"], p["flags"], BASE, DOC["resourceFlavors"])
           for k, (_, _, _, p) in enumerate(sc)]
    return cpp_host_gen.build_driver(cpp, WORK), sc


def test_cpp_flatten_matches_python(driver):
    exe, sc = driver
    got = json.loads(subprocess.run([exe, "flatten"], check=True, capture_output=True, text=True).stdout)
    assert len(got) == len(sc)
    for g, (name, snap, idx, _) in zip(got, sc):
        assert g["cqs"] == idx.cqs and g["cohorts"] == idx.cohorts and g["flavors"] == idx.flavors and g["resources"] == idx.resources, name
        assert g["pods_resource"] == snap.pods_resource and g["flags"] == snap.flags and g["now_ns"] == snap.now_ns, name
        a = snap.arrays
        for k, v in g.items():
            if k in ("cqs", "cohorts", "flavors", "resources", "pods_resource", "flags", "now_ns", "adm_use_fr", "adm_use_qty"):
                continue
            want = np.asarray(a[k]).reshape(-1)
            assert np.array_equal(np.asarray(v, dtype=want.dtype if k != "fair_weight" else np.float64), want), (name, k)
        # usage cells of an admitted workload are an unordered map
        st = g["adm_use_start"]
        for i in range(len(st) - 1):
            gc = sorted(zip(g["adm_use_fr"][st[i]:st[i + 1]], g["adm_use_qty"][st[i]:st[i + 1]]))
            wc = sorted(zip(a["adm_use_fr"][st[i]:st[i + 1]].tolist(), a["adm_use_qty"][st[i]:st[i + 1]].tolist()))
            assert gc == wc, (name, "usage of admitted", i)


@pytest.mark.gpu
def test_cpp_scheduler_matches_oracle(driver):
    import oracle
    exe, sc = driver
    got = json.loads(subprocess.run([exe, "schedule"], check=True, capture_output=True, text=True).stdout)
    for g, (name, snap, idx, _) in zip(got, sc):
        want = oracle.run_cycle(snap)
        R = snap.n_resource
        assert len(g["entries"]) == snap.n_heads, name
        for e, en in enumerate(g["entries"]):
            assert en["key"] == idx.pending[e], name
            assert (en["decision"], en["mode"], en["borrowing"], en["commitRank"]) == (int(want.decision[e]), int(want.mode[e]), int(want.borrow[e]), int(want.commit_rank[e])), (name, en)
            rows = range(int(snap.arrays["wl_ps_start"][e]), int(snap.arrays["wl_ps_start"][e + 1]))
            for ps, row in zip(en["podSets"], rows):
                assert ps["count"] == int(want.ps_count[row]), (name, en)
                assert ps["flavors"] == {idx.resources[r]: idx.flavors[int(want.ps_flavor[row, r])] for r in range(R) if want.ps_flavor[row, r] >= 0}, (name, en)
            assert en["targets"] == [[idx.admitted[a], rs] for a, rs in want.targets(e)], (name, en)
