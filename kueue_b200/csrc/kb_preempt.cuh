// kb_preempt.cuh — fair-sharing preemption target search on the device (classical / hierarchical: kb_search.cuh).
//
// Reference: pkg/scheduler/preemption/preemption.go:127-153,238-314,547-584,
// preemption_oracle.go:41-71, classical/candidate_generator.go:52-162,
// classical/hierarchical_preemption.go:72-227, common/ordering.go:41-100,
// common/preemption_policy.go:30-48, pkg/cache/scheduler/resource_node.go:223-255.
//
// The search mutates the quota tree (remove candidate / fill back / restore), so every CTA
// works on a PRIVATE copy of the preemptor's root tree: in shared memory when the tree fits,
// else in a per-CTA global scratch.  One search is a data-dependent sequential walk (classify,
// remove, re-check, fill back), so it runs on ONE thread; parallelism comes from running many
// searches at once (one single-warp CTA each; the other lanes only stage the tree).  The
// per-ClusterQueue candidate queues are cursors into the ClusterQueue's list of admitted
// workloads, which is ranked once per cycle by the preemptor-independent keys of
// CandidatesOrdering — no gathering and no sort per preemptor.  (A barrier-based master/worker
// split inside one warp is not expressible: bar.sync is warp-aligned.)
#pragma once

#include "kb_device.cuh"

// Private, mutable view of one root tree; node handle = local index inside the tree.
template <bool kSmem>
struct PTab {
  const DevSnap *D;
  const int32_t *nodes;  // local -> global node id
  int nn, FR;
  i64 *usage;            // [nn][FR] private copy (smem or global scratch); [nn] in column mode
  int col_fr = -1;       // column mode: the view holds ONE flavor-resource column (single-cell searches of the
                         // preemption oracle only ever read and write that column: columns are independent)
  const i64 *sub, *lq, *bl;  // smem copies (kSmem) — unused otherwise
  const int *lparent;        // smem (kSmem) — unused otherwise
  uint8_t *dirty = nullptr;  // fair search: per node, "usage changed since its DominantResourceShare was cached"
  double *drs_ratio = nullptr; int8_t *drs_meta = nullptr;  // cached DRS per node: unweighted ratio; bit 7 borrowing, low bits dominant resource + 1
  __device__ __forceinline__ i64 U(int h, int fr) const { return usage[col_fr >= 0 ? h : h * FR + fr]; }
  __device__ __forceinline__ void setU(int h, int fr, i64 v) const { usage[col_fr >= 0 ? h : h * FR + fr] = v; }
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
      if (dirty) dirty[h] = 1;
      int p = parent(h);
      if (p < 0 || !(val > la)) break;
      val -= la; h = p;
    }
  }
  __device__ inline void remove(int h, int fr, i64 val) const {  // removeUsage :149-158
    while (true) {
      i64 u = U(h, fr), stored = u - LQ(h, fr);
      setU(h, fr, u - val);
      if (dirty) dirty[h] = 1;
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
    for (int k = D->adm_use_start[a]; k < D->adm_use_start[a + 1]; k++) { int fr = D->adm_use_fr[k]; if (col_fr < 0 || fr == col_fr) remove(h, fr, D->adm_use_qty[k]); }
  }
  __device__ inline void add_adm(int a) const {  // Snapshot.AddWorkload :59-64
    int h = handle(D->adm_cq[a]);
    for (int k = D->adm_use_start[a]; k < D->adm_use_start[a + 1]; k++) { int fr = D->adm_use_fr[k]; if (col_fr < 0 || fr == col_fr) add(h, fr, D->adm_use_qty[k]); }
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
  int overflow;  // the candidate list did not fit the scratch of this searcher (speculative lane searches only)
};

// View of a scratch array of one searcher (contiguous; an interleaved layout for the 32 lane searchers of a CTA
// was tried and measured slower, see k_nominate_search).
template <typename T>
struct SArr {
  T *p;
  __device__ __forceinline__ T &operator[](int i) const { return p[i]; }
};

// global scratch of one searcher
struct PreScratch {
  SArr<int32_t> cand;     // ordered candidate list
  SArr<uint8_t> variant;  // preemptionVariant per candidate
  SArr<int32_t> tgt;      // targets of the current search (adm index)
  SArr<uint8_t> tgt_reason;
  SArr<int8_t> cq_class;  // per tree node (handle): 0 none, 1 hierarchy candidates, 2 priority candidates
  SArr<int8_t> on_path;   // per tree node (handle): level on the preemptor's path or -1
  SArr<int32_t> cq_lca;   // per tree node (handle): handle of the subtree root that collected it
  SArr<int32_t> aux1, aux2;  // [adm cap] sort keys / fair sharing: next-in-queue links, retry candidates
  int cap;                // capacity of the [adm cap] arrays
};

__device__ __forceinline__ bool satisfies_policy(const DevSnap &D, const PreCtx &c, int a, int policy) {  // preemption_policy.go:30-48
  int cp = D.adm_priority[a];
  bool lower = c.prio > cp;
  if (policy == KB_POLICY_LOWER_PRIORITY) return lower;
  if (policy == KB_POLICY_LOWER_OR_NEWER_EQUAL_PRIORITY) {
    bool newer = c.prio == cp && c.ts < D.adm_ts[a];
    if (newer && (D.flags & KB_F_TS_PREEMPTION_BUFFER)) newer = D.adm_ts[a] - c.ts > 300ll * 1000000000ll;  // timestampPreemptionBuffer :28
    return lower || newer;
  }
  return policy == KB_POLICY_ANY;
}
__device__ __forceinline__ bool uses_resources(const DevSnap &D, const PreCtx &c, int a) {  // WorkloadUsesResources candidate_generator.go:52-61
  for (int k = D.adm_use_start[a]; k < D.adm_use_start[a + 1]; k++) {
    int fr = D.adm_use_fr[k];
    for (int j = 0; j < c.n_need; j++) if (c.need_fr[j] == fr) return true;
  }
  return false;
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

// ---------------------------------------------------------------------------
// Fair-sharing preemption (preemption.go:338-478, fairsharing/ordering.go:46-208,
// fairsharing/target.go, least_common_ancestor.go, strategy.go).
// ---------------------------------------------------------------------------
// CandidatesOrdering common/ordering.go:41-100 (AdmissionFairSharing keys not modelled)
__device__ inline int cand_ordering(const DevSnap &D, int a, int b, int cq) {
  int ea = D.adm_evicted[a], eb = D.adm_evicted[b];
  if (ea != eb) return ea ? -1 : 1;
  bool ain = D.adm_cq[a] == cq, bin = D.adm_cq[b] == cq;
  if (ain != bin) return bin ? -1 : 1;
  int pa = D.adm_priority[a], pb = D.adm_priority[b];
  if (pa != pb) return pa < pb ? -1 : 1;
  i64 ta = D.adm_qr_ts[a] == INT64_MIN ? D.now_ns : D.adm_qr_ts[a], tb = D.adm_qr_ts[b] == INT64_MIN ? D.now_ns : D.adm_qr_ts[b];
  if (ta != tb) return tb < ta ? -1 : 1;
  i64 ua = D.adm_uid[a], ub = D.adm_uid[b];
  if (ua != ub) return ua < ub ? -1 : 1;
  return 0;
}
template <bool kSmem>
__device__ inline DevDRS fair_drs_compute(const DevSnap &D, const PTab<kSmem> &T, int h, DevDRS d, int p);
// dominantResourceShare fair_sharing.go:126-156 on the private tree.  The share of a node only changes when the usage
// of the node changes (addUsage / removeUsage mark every node they touch), so it is cached per node: the tournament of
// nextTarget (ordering.go:141-208) re-reads the shares of ALL children of a cohort after every popped candidate, but
// only the nodes on that candidate's path were modified.
template <bool kSmem>
__device__ inline DevDRS fair_drs(const DevSnap &D, const PTab<kSmem> &T, int h) {
  int node = T.nodes[h];
  DevDRS d{D.fair_weight[node], 0.0, -1, false};
  int p = D.parent[node];
  if (p < 0) return d;
  if (T.dirty && !T.dirty[h]) {
    int8_t m = T.drs_meta[h];
    d.ratio = T.drs_ratio[h]; d.borrowing = m & 0x40; d.res = (m & 0x3f) - 1;
    return d;
  }
  d = fair_drs_compute(D, T, h, d, p);
  if (T.dirty) { T.dirty[h] = 0; T.drs_ratio[h] = d.ratio; T.drs_meta[h] = (int8_t)((d.borrowing ? 0x40 : 0) | ((d.res + 1) & 0x3f)); }
  return d;
}
template <bool kSmem>
__device__ inline DevDRS fair_drs_compute(const DevSnap &D, const PTab<kSmem> &T, int h, DevDRS d, int p) {
  const int R = D.R, F = D.F;
  for (int r = 0; r < R; r++) {
    i64 b = 0;
    for (int f = 0; f < F; f++) {
      int fr = f * R + r;
      i64 over = T.U(h, fr) - T.Sub(h, fr);
      if (over > 0) b += over;
    }
    if (b > 0) {
      const i64 lend = D.fs_lend[(size_t)p * R + r];  // sum over flavors of the parent's potentialAvailable (k_fair_prep)
      d.borrowing = true;
      if (lend > 0) {
        double ratio = (double)b * 1000.0 / (double)lend;
        if (ratio > d.ratio) { d.ratio = ratio; d.res = r; }
      }
    }
  }
  return d;
}

// fairPreemptions preemption.go:433-478.  Called by ONE thread.
template <bool kSmem>
__device__ inline void fair_search(const DevSnap &D, const PTab<kSmem> &T, PreCtx *c, const PreScratch &S) {
  const int cq = c->cq;
  const int hcq = T.handle(cq);
  const bool has_parent = D.parent[cq] >= 0;
  const bool cohort_cands = has_parent && D.cq_reclaim_within[cq] != KB_POLICY_NEVER;
  const bool own_cands = D.cq_within_cq[cq] != KB_POLICY_NEVER;
  {
    int pl = 0;
    for (int t = hcq; t >= 0; t = T.parent(t)) c->path[pl++] = t;
    c->plen = pl;
    c->n_targets = 0;
    c->overflow = 0;
  }
  for (int h = 0; h < T.nn; h++) { S.on_path[h] = -1; S.cq_class[h] = 0; S.cq_lca[h] = -1; if (T.dirty) T.dirty[h] = 1; }
  for (int k = 1; k < c->plen; k++) S.on_path[c->path[k]] = (int8_t)k;  // preemptorAncestors
  // ---- findCandidates :514-533 + MakeClusterQueueOrdering ordering.go:62-83.  The reference sorts all candidates by
  // CandidatesOrdering and splits them into one queue per ClusterQueue; only the order INSIDE a queue is ever used
  // (nextTarget compares queue heads, pop takes a head).  Inside one ClusterQueue that order is (evicted first,
  // priority, reservation time, UID) = the preemptor-independent rank the per-ClusterQueue lists (cq_adm) are kept
  // in.  So a queue is a cursor into the ClusterQueue's rank-ordered list that skips the workloads the preemptor may
  // not preempt: no gathering, no sort.  Other ClusterQueues qualify only while borrowing (cqIsBorrowing :535-545)
  // -> subset of k_over's list.
  const int slot = D.root_slot[cq];
  const int32_t *over = D.over_list + D.root_cq_start[slot];
  const int n_over = cohort_cands ? D.over_count[slot] : 0;
  auto is_borrowing = [&](int q) {
    int h = T.handle(q);
    for (int j = 0; j < c->n_need; j++) if (T.borrowing_with(h, c->need_fr[j], 0)) return true;
    return false;
  };
  SArr<int32_t> head = S.cq_lca;   // per node: position of the queue head (in cq_adm, or in the retry list), or -1
  SArr<int32_t> next = S.aux1;     // retry list: next of the same ClusterQueue
  SArr<int8_t> pruned = S.cq_class;
  SArr<int32_t> retry = S.aux2;
  bool second = false;             // queues over the retry list (runSecondFsStrategy)
  auto advance = [&](int q, int i) -> int {  // first candidate of ClusterQueue q at or after position i of cq_adm
    const int policy = q == cq ? D.cq_within_cq[cq] : D.cq_reclaim_within[cq];
    for (const int e = D.cq_adm_start[q + 1]; i < e; i++) {
      int a = D.cq_adm[i];
      if (satisfies_policy(D, *c, a, policy) && uses_resources(D, *c, a)) return i;
    }
    return -1;
  };
  auto front = [&](int h) -> int { int i = head[h]; return second ? retry[i] : D.cq_adm[i]; };  // head[h] >= 0
  auto pop = [&](int h) -> int {
    int i = head[h], a;
    if (second) { a = retry[i]; head[h] = next[i]; }
    else { a = D.cq_adm[i]; head[h] = advance(T.nodes[h], i + 1); }
    return a;
  };
  const int n_root_adm = D.root_adm_start[slot + 1] - D.root_adm_start[slot];  // bound on the candidates (loop guards)
  int nall = 0;  // ClusterQueues with at least one candidate
  if (own_cands) { int i = advance(cq, D.cq_adm_start[cq]); head[hcq] = i; nall += i >= 0; }
  for (int k = 0; k < n_over; k++) {
    int q = over[k];
    if (q == cq || !is_borrowing(q)) continue;
    int i = advance(q, D.cq_adm_start[q]);
    head[T.handle(q)] = i; nall += i >= 0;
  }
  if (nall == 0) { c->n_targets = 0; return; }
  auto build_retry_queues = [&](int n) {  // MakeClusterQueueOrdering over the retry candidates
    for (int h = 0; h < T.nn; h++) { head[h] = -1; pruned[h] = 0; }
    for (int i = n - 1; i >= 0; i--) { int h = T.handle(D.adm_cq[retry[i]]); next[i] = head[h]; head[h] = i; }
    second = true;
  };
  auto usage_add = [&](bool add) {  // SimulateUsageAddition / Removal of the incoming workload
    for (int j = 0; j < c->n_use; j++) { if (add) T.add(hcq, c->use_fr[j], c->use_q[j]); else T.remove(hcq, c->use_fr[j], c->use_q[j]); }
  };
  auto fits_fs = [&]() {  // workloadFitsForFairSharing :567-572
    usage_add(false);
    bool r = workload_fits(T, *c, hcq, true);
    usage_add(true);
    return r;
  };
  auto next_target = [&](int root) -> int {  // nextTarget ordering.go:141-208 (tail recursion unrolled)
    int cohort = root;
    for (int guard = 0;; guard++) {
      if (guard > T.nn + 2) { atomicOr(D.status, KBS_INTERNAL_LOOP); pruned[root] = 1; return -1; }
      int cnode = T.nodes[cohort];
      int highest_cq = -1; DevDRS hcq_drs{1.0, -1.0, -1, false};
      int highest_co = -1; DevDRS hco_drs{1.0, -1.0, -1, false};
      for (int k = D.child_start[cnode]; k < D.child_start[cnode + 1]; k++) {
        int ch = D.child_list[k];
        if (ch >= D.Q) continue;
        int h = T.handle(ch);
        if (pruned[h]) continue;
        DevDRS drs = fair_drs(D, T, h);
        if ((!drs.borrowing && h != hcq) || head[h] < 0) pruned[h] = 1;
        else {
          int cmp = drs_compare(drs, hcq_drs);
          if (cmp == 0) {
            if (cand_ordering(D, front(h), front(highest_cq), cq) < 0) highest_cq = h;
          } else if (cmp == 1) { hcq_drs = drs; highest_cq = h; }
        }
      }
      for (int k = D.child_start[cnode]; k < D.child_start[cnode + 1]; k++) {
        int ch = D.child_list[k];
        if (ch < D.Q) continue;
        int h = T.handle(ch);
        if (pruned[h]) continue;
        DevDRS drs = fair_drs(D, T, h);
        if (!drs.borrowing && S.on_path[h] < 0) pruned[h] = 1;
        else if (drs_compare(drs, hco_drs) >= 0) { hco_drs = drs; highest_co = h; }
      }
      if (highest_co < 0 && highest_cq < 0) { pruned[cohort] = 1; return -1; }
      if (drs_compare(hco_drs, hcq_drs) >= 0) { cohort = highest_co; continue; }
      return highest_cq;
    }
  };
  auto almost_lcas = [&](int htarget, int *pre_al, int *tgt_al) {  // least_common_ancestor.go:27-58
    int lca = -1;
    for (int t = T.parent(htarget); t >= 0; t = T.parent(t)) if (S.on_path[t] >= 0) { lca = t; break; }
    auto al = [&](int h) { int a = h; for (int t = T.parent(h); t >= 0; t = T.parent(t)) { if (t == lca) return a; a = t; } return a; };
    *pre_al = al(hcq); *tgt_al = al(htarget);
  };
  // parseStrategies :319-333
  bool s2a = D.flags & KB_F_FS_STRATEGY_S2A, s2b = D.flags & KB_F_FS_STRATEGY_S2B;
  int strat[2], nstrat;
  if (!s2a && !s2b) { strat[0] = 0; strat[1] = 1; nstrat = 2; }
  else if (s2a && s2b) { if (D.flags & KB_F_FS_STRATEGY_S2B_FIRST) { strat[0] = 1; strat[1] = 0; } else { strat[0] = 0; strat[1] = 1; } nstrat = 2; }
  else { strat[0] = s2a ? 0 : 1; nstrat = 1; }

  usage_add(true);  // :446 DRS values must include the incoming workload
  int nt = 0, nretry = 0;
  bool fits = false;
  {  // runFirstFsStrategy :338-403
    bool within_nominal = false;
    if (D.flags & KB_F_FS_PREEMPT_WITHIN_NOMINAL) {  // queueWithinNominalInResourcesNeedingPreemption :591-598
      within_nominal = true;
      for (int j = 0; j < c->n_need; j++) if (T.borrowing_with(hcq, c->need_fr[j], 0)) within_nominal = false;
    }
    auto take = [&](int a, int reason) { T.remove_adm(a); S.tgt[nt] = a; S.tgt_reason[nt] = (uint8_t)reason; nt++; };
    auto step = [&](int h) -> bool {
      if (h == hcq) { take(pop(h), KB_REASON_IN_CLUSTER_QUEUE); return fits_fs(); }
      if (within_nominal) { take(pop(h), KB_REASON_IN_COHORT_RECLAMATION); return fits_fs(); }
      int pa, ta; almost_lcas(h, &pa, &ta);
      DevDRS pre_new = fair_drs(D, T, pa), tgt_old = fair_drs(D, T, ta);  // ComputeShares target.go:53-56
      while (head[h] >= 0) {
        int a = pop(h);
        T.remove_adm(a);  // ComputeTargetShareAfterRemoval :66-73
        int p2, t2; almost_lcas(h, &p2, &t2);
        DevDRS tgt_new = fair_drs(D, T, t2);
        T.add_adm(a);
        bool ok = strat[0] == 0 ? drs_compare(pre_new, tgt_new) <= 0 : drs_compare(pre_new, tgt_old) < 0;  // strategy.go:41-48
        if (ok) { take(a, KB_REASON_IN_COHORT_FAIR_SHARING); if (fits_fs()) return true; break; }
        retry[nretry++] = a;
      }
      return false;
    };
    if (!has_parent) {
      while (head[hcq] >= 0) if (step(hcq)) { fits = true; break; }
    } else {
      int root = c->path[c->plen - 1];
      for (int guard = 0; !pruned[root]; guard++) {
        if (guard > 4 * (T.nn + n_root_adm) + 64) { atomicOr(D.status, KBS_INTERNAL_LOOP); break; }
        int h = next_target(root);
        if (h < 0) continue;
        if (step(h)) { fits = true; break; }
      }
    }
  }
  if (!fits && nstrat > 1 && has_parent) {  // runSecondFsStrategy :407-431
    build_retry_queues(nretry);
    int root = c->path[c->plen - 1];
    for (int guard = 0; !pruned[root]; guard++) {
      if (guard > 4 * (T.nn + n_root_adm) + 64) { atomicOr(D.status, KBS_INTERNAL_LOOP); break; }
      int h = next_target(root);
      if (h < 0) continue;
      int pa, ta; almost_lcas(h, &pa, &ta);
      DevDRS pre_new = fair_drs(D, T, pa), tgt_old = fair_drs(D, T, ta);
      if (drs_compare(pre_new, tgt_old) < 0) {
        int a = pop(h);
        T.remove_adm(a); S.tgt[nt] = a; S.tgt_reason[nt] = KB_REASON_IN_COHORT_FAIR_SHARING; nt++;
        if (fits_fs()) { fits = true; break; }
      }
      pruned[h] = 1;  // DropQueue
    }
  }
  usage_add(false);  // revertSimulation :459
  if (fits) {
    for (int k = nt - 2; k >= 0; k--) {  // fillBackWorkloads(allowBorrowing = true) :469
      T.add_adm(S.tgt[k]);
      if (workload_fits(T, *c, hcq, true)) { S.tgt[k] = S.tgt[nt - 1]; S.tgt_reason[k] = S.tgt_reason[nt - 1]; nt--; }
      else T.remove_adm(S.tgt[k]);
    }
  }
  for (int k = 0; k < nt; k++) T.add_adm(S.tgt[k]);  // restoreSnapshot
  c->n_targets = fits ? nt : 0;
}

