"""Seeded synthetic snapshots for the BASELINE.json configs.

Generated directly in the flat SoA form with numpy (PCG64, seed = config index)
so 1M-workload snapshots build in about a second.  Value ranges keep every sum far
below 2^62.
"""
from __future__ import annotations

import numpy as np

from . import abi


def _cohort_forest(n_root: int, fanouts, cqs_per_leaf: int):
    """Cohort forest: n_root roots, then `fanouts` children per level; CQs hang off
    the deepest cohorts.  Returns (parent_of_cohort, leaf cohort ids)."""
    parents = [-1] * n_root
    level = list(range(n_root))
    for f in fanouts:
        nxt = []
        for p in level:
            for _ in range(f):
                parents.append(p)
                nxt.append(len(parents) - 1)
        level = nxt
    return np.array(parents, np.int64), np.array(level, np.int64)


def make_snapshot(config: int = 2, W: int | None = None, Q: int | None = None, F: int | None = None,
                  R: int | None = None, seed: int | None = None, heads: str = "all",
                  podsets_max: int = 1, admitted: int | None = None, preemption: bool | None = None,
                  partial: bool = False, tight: float = 1.0) -> abi.FlatSnapshot:
    """config 1: 100 wl x 10 CQ x 2 flavors x 3 resources, no cohort
       config 2: 100k x 1k x 8 x 4, StrictFIFO, no cohort, usage 60-90% of nominal
       config 3: 1M x 10k, 100 flat cohorts x 100 CQ, BestEffortFIFO, fair sharing
       config 4: 1M x 10k, depth-4 cohort forest (10 -> 5 -> 5 -> 4), classical
    heads: "all" (batched evaluator over every pending workload) | "one_per_cq"
    (reference cycle: the best-ordered workload of each CQ)."""
    dflt = {1: (100, 10, 2, 3), 2: (100_000, 1_000, 8, 4), 3: (1_000_000, 10_000, 8, 4), 4: (1_000_000, 10_000, 8, 4)}[config]
    W = W or dflt[0]; Q = Q or dflt[1]; F = F or dflt[2]; R = R or dflt[3]
    rng = np.random.Generator(np.random.PCG64(config if seed is None else seed))
    FR = F * R
    # ---- cohort structure ----
    if config in (1, 2):
        coh_parent = np.zeros(0, np.int64); cq_parent = np.full(Q, -1, np.int64)
    elif config == 3:
        ncoh = max(1, Q // 100)
        coh_parent = np.full(ncoh, -1, np.int64)
        cq_parent = Q + (np.arange(Q) % ncoh)
    else:
        nroot = max(1, Q // 1000)
        coh_parent, leaves = _cohort_forest(nroot, (5, 5, 4), 10)
        cq_parent = Q + leaves[np.arange(Q) % len(leaves)]
        coh_parent = np.where(coh_parent >= 0, coh_parent + Q, -1)
    C = len(coh_parent); N = Q + C
    snap = abi.FlatSnapshot(n_cq=Q, n_cohort=C, n_flavor=F, n_resource=R)
    flags = abi.FLAGS_DEFAULT
    if config == 3:
        flags |= abi.F_FAIR_SHARING
    snap.flags = flags
    snap.set("parent", np.concatenate([cq_parent, coh_parent]))
    fw = np.ones(N)
    if config == 3:
        fw[:Q] = rng.choice([0.5, 1.0, 2.0], Q)
    snap.set("fair_weight", fw)
    # ---- quotas ----
    scale = np.array([2_000, 2**31, 4, 100_000, 50, 10, 1000, 64][:R], np.int64)  # mean request per resource
    per_cq = max(1, W // Q)
    k = rng.uniform(4, 16, (Q, F, 1))
    nominal = np.zeros((N, FR), np.int64)
    nominal[:Q] = (k * scale[None, None, :]).astype(np.int64).reshape(Q, FR)
    bl = np.full((N, FR), abi.KB_NO_LIMIT, np.int64)
    ll = np.full((N, FR), abi.KB_NO_LIMIT, np.int64)
    if config == 3:
        lend = rng.random(Q) < 0.2
        ll[:Q][lend] = nominal[:Q][lend] // 2
    if config == 4:
        has_q = rng.random(C) < 0.3
        nominal[Q:][has_q] = (rng.uniform(2, 8, (int(has_q.sum()), F, 1)) * scale[None, None, :]).astype(np.int64).reshape(-1, FR)
        blm = rng.random(Q) < 0.3
        bl[:Q][blm] = nominal[:Q][blm]
    snap.set("nominal", nominal); snap.set("borrow_limit", bl); snap.set("lend_limit", ll)
    lo, hi = {1: (0.6, 0.9), 2: (0.6, 0.9), 3: (0.7, 1.25), 4: (0.85, 1.2)}[config]
    lo, hi = lo * tight, hi * tight
    if preemption is None:
        preemption = config == 4
    if admitted is None:
        admitted = max(W // 5, 2 * Q * F) if preemption else 0
    if admitted:
        # admitted workloads (preemption candidates); ClusterQueue usage is exactly their sum
        A = admitted
        pair = rng.permutation(A) % (Q * F)        # round-robin over (ClusterQueue, flavor) pairs
        a_cq, a_f = pair // F, pair % F
        per = np.bincount(pair, minlength=Q * F)[pair].astype(np.float64)
        frac = rng.uniform(lo, hi, Q * F)[pair]      # target usage of the pair as a fraction of nominal
        qty = (nominal[:Q].reshape(Q, F, R)[a_cq, a_f] * (frac / per)[:, None] * rng.uniform(0.7, 1.3, (A, R))).astype(np.int64)
        usage = np.zeros((Q, F, R), np.int64)
        np.add.at(usage, (a_cq, a_f), qty)
        usage = usage.reshape(Q, FR)
        snap.set("adm_cq", a_cq); snap.set("adm_priority", rng.integers(0, 4, A) * 100)
        snap.set("adm_ts", 1_600_000_000_000_000_000 + rng.permutation(A).astype(np.int64) * 1_000_000)
        qr = 1_650_000_000_000_000_000 + rng.permutation(A).astype(np.int64) * 1_000_000
        qr[rng.random(A) < 0.02] = abi.KB_TS_UNSET
        snap.set("adm_qr_ts", qr); snap.set("adm_uid", rng.permutation(A) + 10_000_000)
        snap.set("adm_evicted", rng.random(A) < 0.01)
        snap.set("adm_use_start", np.arange(A + 1) * R)
        snap.set("adm_use_fr", (a_f[:, None] * R + np.arange(R)[None, :]).reshape(-1))
        snap.set("adm_use_qty", qty.reshape(-1))
        snap.now_ns = 1_700_000_000_000_000_000
    else:
        usage = (nominal[:Q] * rng.uniform(lo, hi, (Q, FR))).astype(np.int64)
        if config in (1, 2):
            usage = np.minimum(usage, nominal[:Q])
    snap.set("cq_usage", usage)
    if preemption:
        snap.set("cq_within_cq", rng.choice([abi.POLICY_NEVER, abi.POLICY_LOWER_PRIORITY, abi.POLICY_LOWER_OR_NEWER_EQUAL_PRIORITY], Q))
        if C:
            snap.set("cq_reclaim_within", rng.choice([abi.POLICY_NEVER, abi.POLICY_LOWER_PRIORITY, abi.POLICY_ANY], Q))
            bw = rng.random(Q) < 0.4
            bw &= snap.arrays["cq_reclaim_within"] != abi.POLICY_NEVER  # API validation: reclaim=Never excludes borrowWithinCohort
            snap.set("cq_borrow_within", np.where(bw, abi.POLICY_LOWER_PRIORITY, abi.POLICY_NEVER))
            thr = bw & (rng.random(Q) < 0.5)
            snap.set("cq_has_bwc_threshold", thr); snap.set("cq_bwc_threshold", np.where(thr, 100, 0))
    # ---- CQ attributes: one resource group covering all resources, all flavors ----
    snap.set("cq_strategy", np.full(Q, abi.QUEUE_STRICT_FIFO if config == 2 else abi.QUEUE_BEST_EFFORT_FIFO))
    snap.set("cq_rg_start", np.arange(Q + 1))
    snap.set("rg_res_mask", np.full(Q, (1 << R) - 1))
    snap.set("rg_flavor_start", np.arange(Q + 1) * F)
    snap.set("rg_flavors", np.tile(np.arange(F), Q))
    # ---- pending workloads, grouped by CQ ----
    wl_cq = np.sort(rng.integers(0, Q, W)) if config != 1 else np.repeat(np.arange(Q), per_cq)[:W]
    snap.set("wl_cq", wl_cq)
    snap.set("wl_priority", rng.integers(0, 4, W) * 100)
    snap.set("wl_ts", 1_700_000_000_000_000_000 + rng.permutation(W).astype(np.int64) * 1_000_000)
    snap.set("wl_uid", rng.permutation(W))
    nps = rng.integers(1, podsets_max + 1, W) if podsets_max > 1 else np.ones(W, np.int64)
    st = np.concatenate([[0], np.cumsum(nps)])
    P = int(st[-1])
    snap.set("wl_ps_start", st)
    count = rng.integers(1, 9, P)
    per_pod = (rng.uniform(0.05, 1.0, (P, R)) * scale[None, :] / 4).astype(np.int64)
    big = rng.random(P) < 0.03  # a few podsets larger than any ClusterQueue can ever hold
    per_pod[big] *= 40
    snap.set("ps_req", per_pod * count[:, None])
    snap.set("ps_req_mask", np.full(P, (1 << R) - 1))
    snap.set("ps_count", count)
    if partial:
        mc = np.where(rng.random(P) < 0.5, np.maximum(1, count // 2), -1)
        snap.set("ps_min_count", mc)
    ok = rng.integers(0, 2**63, P, dtype=np.uint64) | rng.integers(0, 2**63, P, dtype=np.uint64) | np.uint64(1 << (F - 1))
    snap.set("ps_flavor_ok", ok)  # ~75% of flavors eligible, last flavor always
    if heads == "all":
        snap.set("heads", np.arange(W))
    else:
        # reference Heads(): per CQ the first in (priority desc, ts asc, uid) order (cluster_queue.go:636-685)
        order = np.lexsort((snap.wl_uid, snap.wl_ts, -snap.wl_priority, snap.wl_cq))
        first = np.concatenate([[True], snap.wl_cq[order][1:] != snap.wl_cq[order][:-1]])
        snap.set("heads", order[first])
    return snap.finalize()


def compact_to_heads(snap: abi.FlatSnapshot) -> abi.FlatSnapshot:
    """The pending tables reduced to the entries of the cycle — what `queues.Heads()` hands the
    reference scheduler (manager.go:721-794): the cycle never sees the workloads deeper in the queues."""
    from .shard import _csr_take
    a = snap.arrays
    R = snap.n_resource
    h = a["heads"].astype(np.int64)
    out = abi.FlatSnapshot(n_cq=snap.n_cq, n_cohort=snap.n_cohort, n_flavor=snap.n_flavor, n_resource=R,
                           pods_resource=snap.pods_resource, flags=snap.flags, now_ns=snap.now_ns)
    for k, v in a.items():
        if not (k.startswith("wl_") or k.startswith("ps_") or k == "heads"):
            out.arrays[k] = v
    for nm in ("wl_cq", "wl_priority", "wl_ts", "wl_uid", "wl_last_gen"):
        out.set(nm, a[nm][h])
    st, rows = _csr_take(a["wl_ps_start"].astype(np.int64), h)
    out.set("wl_ps_start", st)
    out.set("ps_req", a["ps_req"].reshape(-1, R)[rows]); out.set("ps_last_tried", a["ps_last_tried"].reshape(-1, R)[rows])
    for nm in ("ps_req_mask", "ps_count", "ps_min_count", "ps_flavor_ok"):
        out.set(nm, a[nm][rows])
    if "ps_group" in a:  # optional tables (kb_snapshot: NULL when absent)
        out.set("ps_group", a["ps_group"][rows])
    for nm in ("wl_has_quota_reservation", "wl_sched_hash"):
        if nm in a:
            out.set(nm, a[nm][h])
    out.set("heads", np.arange(len(h)))
    return out.finalize()
