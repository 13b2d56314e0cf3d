"""Second, independent restatement (pure Python, recursive like resource_node.go:104-133,183-217) of the quota-tree
arithmetic, run against the C++ oracle on random hierarchical snapshots."""
import numpy as np
import pytest

import oracle
from kueue_b200 import abi, synth

NO = abi.KB_NO_LIMIT


def _py_tree(snap):
    a = snap.arrays
    N, Q, FR = snap.n_nodes, snap.n_cq, snap.n_fr
    parent = a["parent"].tolist()
    nominal = a["nominal"].reshape(N, FR).astype(object)
    bl = a["borrow_limit"].reshape(N, FR).astype(object)
    ll = a["lend_limit"].reshape(N, FR).astype(object)
    children = [[] for _ in range(N)]
    for n, p in enumerate(parent):
        if p >= 0:
            children[p].append(n)
    sub = [[0] * FR for _ in range(N)]
    use = [[0] * FR for _ in range(N)]

    def local_quota(n, fr):  # :66-71
        return max(0, sub[n][fr] - ll[n][fr]) if ll[n][fr] != NO else 0

    def build(n):  # updateCohortResourceNode / accumulateFromChild :183-217
        for fr in range(FR):
            sub[n][fr] = int(nominal[n][fr])
            use[n][fr] = int(a["cq_usage"].reshape(Q, FR)[n, fr]) if n < Q else 0
        for c in children[n]:
            build(c)
            for fr in range(FR):
                sub[n][fr] += sub[c][fr] - local_quota(c, fr)
                use[n][fr] += max(0, use[c][fr] - local_quota(c, fr))

    def available(n, fr):  # :104-118
        if parent[n] < 0:
            return sub[n][fr] - use[n][fr]
        lq = local_quota(n, fr)
        local = max(0, lq - use[n][fr])
        pa = available(parent[n], fr)
        if bl[n][fr] != NO:
            pa = min((sub[n][fr] - lq) - max(0, use[n][fr] - lq) + bl[n][fr], pa)
        return local + pa

    def potential(n, fr):  # :122-133
        if parent[n] < 0:
            return sub[n][fr]
        pot = local_quota(n, fr) + potential(parent[n], fr)
        if bl[n][fr] != NO:
            pot = min(sub[n][fr] + bl[n][fr], pot)
        return pot

    for n in range(N):
        if parent[n] < 0:
            build(n)
    av = np.array([[max(0, available(q, fr)) for fr in range(FR)] for q in range(Q)], dtype=np.int64)
    po = np.array([[potential(q, fr) for fr in range(FR)] for q in range(Q)], dtype=np.int64)
    return np.array(sub, dtype=np.int64), np.array(use, dtype=np.int64), av, po


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("config", [3, 4])
def test_tree_matches_python_restatement(config, seed):
    snap = synth.make_snapshot(config, W=60, Q=30 if config == 3 else 40, seed=100 + seed, heads="one_per_cq")
    out = oracle.tree_eval(snap)
    sub, use, av, po = _py_tree(snap)
    assert np.array_equal(out.subtree_quota, sub)
    assert np.array_equal(out.usage, use)
    assert np.array_equal(out.available[:snap.n_cq], av)
    assert np.array_equal(out.potential_available[:snap.n_cq], po)
