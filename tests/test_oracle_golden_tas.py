"""Pins the TAS oracle (oracle/kueue_oracle_tas.cpp) to TestFindTopologyAssignments
(pkg/cache/scheduler/tas_cache_test.go:55): expected leaf assignments / failures per podset."""
import pytest

import oracle
from tests.tas_golden import DOC, build, check


@pytest.mark.parametrize("name", list(DOC["cases"]))
def test_find_topology_assignments(name):
    tc = DOC["cases"][name]
    topo, reqs = build(tc)
    out = oracle.tas_find(topo, reqs)
    check(tc, out, topo)
