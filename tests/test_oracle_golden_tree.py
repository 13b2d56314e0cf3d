"""Pins the oracle's resource-node arithmetic to the reference's TestAvailable
(pkg/cache/scheduler/resource_test.go:31) golden values."""
import pytest

import oracle
from kueue_b200.api import flatten
from tests.golden.tree_cases import TREE_CASES


@pytest.mark.parametrize("name", list(TREE_CASES))
def test_available(name):
    tc = TREE_CASES[name]
    # before adding usage: available == potentiallyAvailable (resource_test.go:355-377)
    snap, idx = flatten(tc["cqs"], tc["cohorts"], usage={})
    out = oracle.tree_eval(snap)
    for cq, m in tc["want_potential"].items():
        for (f, r), v in m.items():
            assert out.available[idx.cqs.index(cq), idx.fr(f, r)] == v, (cq, f, r)
            assert out.potential_available[idx.cqs.index(cq), idx.fr(f, r)] == v, (cq, f, r)
    # after AddUsage (resource_test.go:380-405)
    snap, idx = flatten(tc["cqs"], tc["cohorts"], usage=tc["usage"])
    out = oracle.tree_eval(snap)
    for cq, m in tc["want_available"].items():
        for (f, r), v in m.items():
            assert out.available[idx.cqs.index(cq), idx.fr(f, r)] == v, (cq, f, r)
    for cq, m in tc["want_potential"].items():
        for (f, r), v in m.items():
            assert out.potential_available[idx.cqs.index(cq), idx.fr(f, r)] == v, (cq, f, r)
