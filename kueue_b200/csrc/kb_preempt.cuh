// kb_preempt.cuh — classical / hierarchical preemption target search on the device.
//
// Reference: pkg/scheduler/preemption/preemption.go:127-153,238-314,547-584,
// preemption_oracle.go:41-71, classical/candidate_generator.go:52-162,
// classical/hierarchical_preemption.go:72-227, common/ordering.go:41-100,
// common/preemption_policy.go:30-48, pkg/cache/scheduler/resource_node.go:223-255.
//
// The search mutates the quota tree (remove candidate / fill back / restore), so every CTA
// works on a PRIVATE copy of the preemptor's root tree: in shared memory when the tree fits,
// else in a per-CTA global scratch.  Parallel parts (all threads of the CTA): classifying
// the ClusterQueues of the tree against the preemptor (which subtree collected them, with or
// without hierarchical advantage) and building the ordered candidate list by stable stream
// compaction of the root's admitted workloads, which are pre-sorted once per cycle by the
// preemptor-independent keys of CandidatesOrdering.  The greedy remove / fill-back walk is
// inherently sequential and runs on thread 0.
#pragma once

#include "kb_device.cuh"

enum { PV_NEVER = 0, PV_WITHIN_CQ = 1, PV_HIER_RECLAIM = 2, PV_RECLAIM_NO_BORROW = 3, PV_RECLAIM_WHILE_BORROW = 4 };

// Private, mutable view of one root tree; node handle = local index inside the tree.
template <bool kSmem>
struct PTab {
  const DevSnap *D;
  const int32_t *nodes;  // local -> global node id
  int nn, FR;
  i64 *usage;            // [nn][FR] private copy (smem or global scratch)
  const i64 *sub, *lq, *bl;  // smem copies (kSmem) — unused otherwise
  const int *lparent;        // smem (kSmem) — unused otherwise
  __device__ __forceinline__ i64 U(int h, int fr) const { return usage[h * FR + fr]; }
  __device__ __forceinline__ void setU(int h, int fr, i64 v) const { usage[h * FR + fr] = v; }
  __device__ __forceinline__ i64 Sub(int h, int fr) const { return kSmem ? sub[h * FR + fr] : D->subtree[(size_t)nodes[h] * FR + fr]; }
  __device__ __forceinline__ i64 LQ(int h, int fr) const {
    if (kSmem) return lq[h * FR + fr];
    size_t c = (size_t)nodes[h] * FR + fr;
    return local_quota(D->subtree[c], D->llimit[c]);
  }
  __device__ __forceinline__ i64 BL(int h, int fr) const { return kSmem ? bl[h * FR + fr] : D->blimit[(size_t)nodes[h] * FR + fr]; }
  __device__ __forceinline__ int parent(int h) const {
    if (kSmem) return lparent[h];
    int p = D->parent[nodes[h]];
    return p < 0 ? -1 : D->local_idx[p];
  }
  __device__ __forceinline__ int handle(int node) const { return D->local_idx[node]; }
  __device__ __forceinline__ i64 local_avail(int h, int fr) const { return imax(0, LQ(h, fr) - U(h, fr)); }  // resource_node.go:91-93
  __device__ inline i64 avail(int h, int fr) const {  // available :104-118, clamped like ClusterQueueSnapshot.Available
    int path[KB_MAX_DEPTH + 1], pl = 0;
    for (int t = h; t >= 0; t = parent(t)) path[pl++] = t;
    int rt = path[pl - 1];
    i64 a = Sub(rt, fr) - U(rt, fr);
    for (int k = pl - 2; k >= 0; k--) {
      int nd = path[k];
      i64 u = U(nd, fr), l = LQ(nd, fr), b = BL(nd, fr);
      i64 pa = a;
      if (b != KB_NO_LIMIT) pa = imin((Sub(nd, fr) - l) - imax(0, u - l) + b, pa);
      a = imax(0, l - u) + pa;
    }
    return imax(0, a);
  }
  __device__ inline void add(int h, int fr, i64 val) const {  // addUsage :137-145
    while (true) {
      i64 u = U(h, fr), la = imax(0, LQ(h, fr) - u);
      setU(h, fr, u + val);
      int p = parent(h);
      if (p < 0 || !(val > la)) break;
      val -= la; h = p;
    }
  }
  __device__ inline void remove(int h, int fr, i64 val) const {  // removeUsage :149-158
    while (true) {
      i64 u = U(h, fr), stored = u - LQ(h, fr);
      setU(h, fr, u - val);
      int p = parent(h);
      if (stored <= 0 || p < 0) break;
      val = imin(val, stored); h = p;
    }
  }
  // a CQ is "borrowing with val": Usage + val > Nominal (== SubtreeQuota for a CQ); cohort: > SubtreeQuota
  __device__ __forceinline__ bool borrowing_with(int h, int fr, i64 val) const { return U(h, fr) + val > Sub(h, fr); }
  // FindHeightOfLowestSubtreeThatFits hierarchical_preemption.go:214-227 on the private tree
  __device__ inline int find_height(int h, int fr, i64 val) const {
    int p = parent(h);
    if (!borrowing_with(h, fr, val) || p < 0) return 0;
    i64 remaining = val - local_avail(h, fr);
    int t = p, last = p;
    while (t >= 0) {
      if (!borrowing_with(t, fr, remaining)) return D->height[nodes[t]];
      remaining -= local_avail(t, fr);
      last = t; t = parent(t);
    }
    return D->height[nodes[last]];
  }
  __device__ inline void remove_adm(int a) const {  // Snapshot.RemoveWorkload snapshot.go:49-55
    int h = handle(D->adm_cq[a]);
    for (int k = D->adm_use_start[a]; k < D->adm_use_start[a + 1]; k++) remove(h, D->adm_use_fr[k], D->adm_use_qty[k]);
  }
  __device__ inline void add_adm(int a) const {  // Snapshot.AddWorkload :59-64
    int h = handle(D->adm_cq[a]);
    for (int k = D->adm_use_start[a]; k < D->adm_use_start[a + 1]; k++) add(h, D->adm_use_fr[k], D->adm_use_qty[k]);
  }
};

// Shared-memory context of one search.
struct PreCtx {
  int cq, prio; i64 ts;  // preemptor (workload.Info of the incoming workload)
  int n_use; int use_fr[KB_MAX_CELLS]; i64 use_q[KB_MAX_CELLS];  // workloadUsage.Quota
  int n_need; int need_fr[KB_MAX_CELLS];                          // frsNeedPreemption
  int plen; int path[KB_MAX_DEPTH + 1]; int adv_at[KB_MAX_DEPTH + 1];
  int seg_count[6];
  int n_all, n_targets;
  int wsum[4];
  int scan_base;
};

// per-CTA global scratch
struct PreScratch {
  int32_t *cand;     // ordered candidate list
  uint8_t *variant;  // preemptionVariant per candidate
  int32_t *tgt;      // targets of the current search (adm index)
  uint8_t *tgt_reason;
  int8_t *cq_class;  // per tree node (handle): 0 none, 1 hierarchy candidates, 2 priority candidates
  int8_t *on_path;   // per tree node (handle): level on the preemptor's path or -1
  int32_t *cq_lca;   // per tree node (handle): handle of the subtree root that collected it
};

__device__ __forceinline__ bool satisfies_policy(const DevSnap &D, const PreCtx &c, int a, int policy) {  // preemption_policy.go:30-48
  int cp = D.adm_priority[a];
  bool lower = c.prio > cp;
  if (policy == KB_POLICY_LOWER_PRIORITY) return lower;
  if (policy == KB_POLICY_LOWER_OR_NEWER_EQUAL_PRIORITY) return lower || (c.prio == cp && c.ts < D.adm_ts[a]);
  return policy == KB_POLICY_ANY;
}
__device__ __forceinline__ bool uses_resources(const DevSnap &D, const PreCtx &c, int a) {  // WorkloadUsesResources candidate_generator.go:52-61
  for (int k = D.adm_use_start[a]; k < D.adm_use_start[a + 1]; k++) {
    int fr = D.adm_use_fr[k];
    for (int j = 0; j < c.n_need; j++) if (c.need_fr[j] == fr) return true;
  }
  return false;
}
// classifyPreemptionVariant hierarchical_preemption.go:82-114
__device__ inline int classify_variant(const DevSnap &D, const PreCtx &c, int a, bool hier_adv) {
  if (!uses_resources(D, c, a)) return PV_NEVER;
  bool same = D.adm_cq[a] == c.cq;
  int policy = same ? D.cq_within_cq[c.cq] : D.cq_reclaim_within[c.cq];
  if (!satisfies_policy(D, c, a, policy)) return PV_NEVER;
  if (same) return PV_WITHIN_CQ;
  if (hier_adv) return PV_HIER_RECLAIM;
  if (D.cq_borrow_within[c.cq] == KB_POLICY_NEVER) return PV_RECLAIM_NO_BORROW;  // IsBorrowingWithinCohortForbidden :72-78
  int cp = D.adm_priority[a];
  bool above;  // isAboveBorrowingThreshold :116-124
  if (cp >= c.prio) above = true;
  else if (!D.cq_has_bwc_threshold[c.cq]) above = false;
  else above = cp > D.cq_bwc_threshold[c.cq];
  return above ? PV_RECLAIM_NO_BORROW : PV_RECLAIM_WHILE_BORROW;
}
__device__ __forceinline__ int variant_reason(int v) {  // PreemptionReason :49-61
  switch (v) {
    case PV_WITHIN_CQ: return KB_REASON_IN_CLUSTER_QUEUE;
    case PV_HIER_RECLAIM: return KB_REASON_IN_COHORT_RECLAMATION;
    case PV_RECLAIM_WHILE_BORROW: return KB_REASON_IN_COHORT_RECLAIM_WHILE_BORROWING;
    case PV_RECLAIM_NO_BORROW: return KB_REASON_IN_COHORT_RECLAMATION;
  }
  return 0;
}

template <bool kSmem>
__device__ __forceinline__ bool within_nominal(const PTab<kSmem> &T, const PreCtx &c, int h) {  // IsWithinNominalInResources resource_node.go:248-255
  for (int j = 0; j < c.n_need; j++) if (T.U(h, c.need_fr[j]) > T.Sub(h, c.need_fr[j])) return false;
  return true;
}
template <bool kSmem>
__device__ inline bool workload_fits(const PTab<kSmem> &T, const PreCtx &c, int hcq, bool allow_borrowing) {  // preemption.go:550-561
  for (int j = 0; j < c.n_use; j++) {
    if (!allow_borrowing && T.borrowing_with(hcq, c.use_fr[j], c.use_q[j])) return false;
    if (c.use_q[j] > T.avail(hcq, c.use_fr[j])) return false;
  }
  return true;
}

// block-wide exclusive scan of one flag per thread (blockDim <= 128); returns the thread's
// offset, *total = number of set flags.  All threads must call it.
__device__ inline int block_flag_scan(PreCtx *c, bool flag, int *total) {
  unsigned b = __ballot_sync(0xffffffffu, flag);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) c->wsum[w] = __popc(b);
  __syncthreads();
  int base = 0, tot = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) { if (i < w) base += c->wsum[i]; tot += c->wsum[i]; }
  __syncthreads();
  *total = tot;
  return base + __popc(b & ((1u << lane) - 1));
}

// classicalPreemptions preemption.go:238-293.  Called by ALL threads of the CTA with the
// context filled in (cq, prio, ts, use_*, need_*).  On return c->n_targets / S.tgt hold
// the targets (0 = none); the private tree is restored.
template <bool kSmem>
__device__ inline void classical_search(const DevSnap &D, const PTab<kSmem> &T, PreCtx *c, const PreScratch &S) {
  const int cq = c->cq;
  const int hcq = T.handle(cq);
  const bool has_parent = D.parent[cq] >= 0;
  const bool cohort_cands = has_parent && D.cq_reclaim_within[cq] != KB_POLICY_NEVER;
  const bool own_cands = D.cq_within_cq[cq] != KB_POLICY_NEVER;
  // ---- 1. preemptor path and hierarchical advantage per level (collectCandidatesForHierarchicalReclaim :151-177)
  if (threadIdx.x == 0) {
    int pl = 0;
    for (int t = hcq; t >= 0; t = T.parent(t)) c->path[pl++] = t;
    c->plen = pl;
    i64 rem[KB_MAX_CELLS];
    for (int j = 0; j < c->n_use; j++) rem[j] = c->use_q[j];
    auto qfiq = [&](int h) {  // QuantitiesFitInQuota resource_node.go:234-244
      bool fits = true;
      for (int j = 0; j < c->n_use; j++) {
        int fr = c->use_fr[j];
        if (T.U(h, fr) + rem[j] > T.Sub(h, fr)) fits = false;
        rem[j] = imax(0, rem[j] - T.local_avail(h, fr));
      }
      return fits;
    };
    bool adv = qfiq(hcq);
    for (int k = 1; k < pl; k++) {
      c->adv_at[k] = adv;
      bool fits = qfiq(c->path[k]);
      adv = adv || fits;
    }
    c->n_targets = 0;
  }
  for (int h = threadIdx.x; h < T.nn; h += blockDim.x) S.on_path[h] = -1;
  __syncthreads();
  for (int k = threadIdx.x; k < c->plen; k += blockDim.x) S.on_path[c->path[k]] = (int8_t)k;
  __syncthreads();
  // ---- 2. which ClusterQueues are collected, and by which subtree (collectCandidatesInSubtree :181-199)
  for (int h = threadIdx.x; h < T.nn; h += blockDim.x) {
    int cls = 0, lca = -1;
    if (cohort_cands && T.nodes[h] < D.Q && h != hcq && !within_nominal(T, *c, h)) {
      bool ok = true;
      int t = T.parent(h);
      while (t >= 0 && S.on_path[t] < 0) {  // cohorts strictly between the CQ and the subtree root
        if (within_nominal(T, *c, t)) { ok = false; break; }
        t = T.parent(t);
      }
      if (ok && t >= 0) { lca = t; cls = c->adv_at[S.on_path[t]] ? 1 : 2; }
    }
    S.cq_class[h] = (int8_t)cls;
    S.cq_lca[h] = lca;
  }
  __syncthreads();
  // ---- 3. ordered candidate list: evicted{hier, prio, same} then non-evicted{hier, prio, same}
  //         (NewCandidateIterator candidate_generator.go:77-121), by stable compaction of the
  //         root's admitted workloads pre-sorted by (evicted, priority asc, newer first, uid).
  int slot = D.root_slot[cq];
  int a0 = D.root_adm_start[slot], a1 = D.root_adm_start[slot + 1];
  int nall = 0;
  for (int seg = 0; seg < 6; seg++) {
    int ev = seg < 3 ? 1 : 0, cls = seg % 3 + 1;
    int seg_n = 0;
    if ((cls == 3 && own_cands) || (cls != 3 && cohort_cands)) {
      for (int base = a0; base < a1; base += blockDim.x) {
        int i = base + threadIdx.x;
        bool flag = false; int a = -1, v = PV_NEVER;
        if (i < a1) {
          a = D.adm_sorted[i];
          if ((int)D.adm_evicted[a] == ev) {
            int acq = D.adm_cq[a];
            int acls = acq == cq ? 3 : S.cq_class[T.handle(acq)];
            if (acls == cls) { v = classify_variant(D, *c, a, cls == 1); flag = v != PV_NEVER; }
          }
        }
        int tot;
        int pos = block_flag_scan(c, flag, &tot);
        if (flag) { S.cand[nall + seg_n + pos] = a; S.variant[nall + seg_n + pos] = (uint8_t)v; }
        seg_n += tot;
      }
    }
    if (threadIdx.x == 0) c->seg_count[seg] = seg_n;
    nall += seg_n;
  }
  __syncthreads();
  // ---- 4. greedy remove / fill back (thread 0)
  if (threadIdx.x == 0) {
    int n_hier = c->seg_count[0] + c->seg_count[3], n_prio = c->seg_count[1] + c->seg_count[4];
    bool no_other = n_hier == 0 && n_prio == 0, no_hier = n_hier == 0;
    bool forbidden = D.cq_borrow_within[cq] == KB_POLICY_NEVER;
    bool under_nominal = true;  // queueUnderNominalInResourcesNeedingPreemption :577-584
    for (int j = 0; j < c->n_need; j++) if (T.U(hcq, c->need_fr[j]) >= T.Sub(hcq, c->need_fr[j])) under_nominal = false;
    bool opts[2]; int nopts;
    if (no_other || (forbidden && !under_nominal)) { opts[0] = true; nopts = 1; }   // :266-267
    else if (forbidden && no_hier) { opts[0] = false; opts[1] = true; nopts = 2; }  // :268-269
    else { opts[0] = true; opts[1] = false; nopts = 2; }                            // :270-271
    int nt = 0; bool found = false;
    for (int oi = 0; oi < nopts && !found; oi++) {
      bool borrow = opts[oi];
      nt = 0;
      for (int i = 0; i < nall; i++) {
        int a = S.cand[i], v = S.variant[i];
        int acq = D.adm_cq[a];
        if (acq != cq) {  // candidateIsValid candidate_generator.go:140-162
          if (borrow && v == PV_RECLAIM_NO_BORROW) continue;
          int h = T.handle(acq);
          if (within_nominal(T, *c, h)) continue;
          bool valid = true;
          int lca = S.cq_lca[h];
          for (int t = T.parent(h); t >= 0 && t != lca; t = T.parent(t))
            if (within_nominal(T, *c, t)) { valid = false; break; }
          if (!valid) continue;
        }
        T.remove_adm(a);
        S.tgt[nt] = a; S.tgt_reason[nt] = (uint8_t)variant_reason(v); nt++;
        if (workload_fits(T, *c, hcq, borrow)) {
          for (int k = nt - 2; k >= 0; k--) {  // fillBackWorkloads :295-308
            T.add_adm(S.tgt[k]);
            if (workload_fits(T, *c, hcq, borrow)) { S.tgt[k] = S.tgt[nt - 1]; S.tgt_reason[k] = S.tgt_reason[nt - 1]; nt--; }
            else T.remove_adm(S.tgt[k]);
          }
          found = true;
          break;
        }
      }
      for (int k = 0; k < nt; k++) T.add_adm(S.tgt[k]);  // restoreSnapshot :310-314
    }
    c->n_targets = found ? nt : 0;
  }
  __syncthreads();
}
