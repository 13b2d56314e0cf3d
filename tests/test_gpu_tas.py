"""kb_tas_find on the device: the reference's TestFindTopologyAssignments cases, and synthetic topologies against the oracle."""
import numpy as np
import pytest

import oracle
from kueue_b200 import tas
from tests.tas_golden import DOC, build, check

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ev():
    from kueue_b200 import native
    e = native.Evaluator(0)
    yield e
    e.close()


def _same(got, want):
    assert np.array_equal(got.status, want.status), (got.status.tolist()[:20], want.status.tolist()[:20])
    assert np.array_equal(got.asg_start, want.asg_start)
    n = int(want.asg_start[-1])
    assert np.array_equal(got.asg_leaf[:n], want.asg_leaf[:n]) and np.array_equal(got.asg_count[:n], want.asg_count[:n])


@pytest.mark.parametrize("name", list(DOC["cases"]))
def test_reference_find_topology_assignments(ev, name):
    """TestFindTopologyAssignments (tas_cache_test.go:55): device == the reference's expected assignment == oracle."""
    tc = DOC["cases"][name]
    topo, reqs = build(tc)
    got = ev.tas_find(topo, reqs)
    check(tc, got, topo)
    _same(got, oracle.tas_find(topo, reqs))


@pytest.mark.parametrize("kw", [dict(blocks=2, racks=4, hosts=8, n=200, chains=True), dict(blocks=3, racks=10, hosts=20, n=500, chains=True, used=0.8),
                                dict(blocks=10, racks=20, hosts=50, n=2000, chains=False), dict(blocks=4, racks=8, hosts=16, n=400, chains=True, max_pods=400)])
def test_synthetic_topology_matches_oracle(ev, kw):
    topo = tas.synth_topology(kw["blocks"], kw["racks"], kw["hosts"], used=kw.get("used", 0.5))
    reqs = tas.synth_requests(topo, kw["n"], chains=kw["chains"], max_pods=kw.get("max_pods", 64))
    cap = int(reqs.count.sum()) + 16
    want = oracle.tas_find(topo, reqs, cap)
    assert (want.status == 0).any() and (want.status != 0).any()
    _same(ev.tas_find(topo, reqs, cap), want)
