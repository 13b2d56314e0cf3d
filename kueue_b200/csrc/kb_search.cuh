// kb_search.cuh — classical / hierarchical preemption target search, warp-cooperative.
//
// Reference: pkg/scheduler/preemption/preemption.go:127-153,238-314,547-584,
// preemption_oracle.go:41-71, classical/candidate_generator.go:52-162,
// classical/hierarchical_preemption.go:72-227, common/ordering.go:41-100,
// common/preemption_policy.go:30-48, pkg/cache/scheduler/resource_node.go:223-255.
//
// One search = one WARP.  What the reference does per search, and how it maps here:
//
//   collect + classify + sort candidates   The candidate order (common/ordering.go:41-100) inside one of the six
//   (candidate_generator.go:77-121)        segments {evicted, not} x {hierarchy, priority, same queue} is
//                                          preemptor-independent, so the admitted workloads are ranked ONCE per
//                                          cycle (adm_sorted) and bucketed per (root, flavor-resource) in that order
//                                          (fr-lists, k_frl_*).  A search streams one bucket with coalesced 32 B
//                                          records, 32 candidates per step, and classifies each candidate from
//                                          registers: the "ClusterQueue / cohort above nominal" tests of
//                                          collectCandidatesInSubtree (hierarchical_preemption.go:181-199) are one
//                                          bit mask per (node, flavor-resource) precomputed by k_columns, the lowest
//                                          common ancestor with the preemptor's path is an Euler-interval test.
//                                          No gather, no sort: a one-byte (segment, variant) code per candidate.
//   greedy remove / fill back              The quota tree is column-separable (every flavor-resource is an
//   (preemption.go:238-308)                independent tree), so a search only needs the columns of the cells the
//                                          workload uses: lane j owns column j, private in shared memory
//                                          (node x 8 B), removeUsage / addUsage / available run lane-parallel and
//                                          workloadFits is one warp vote.
//
// The preemption oracle's single-cell calls (SimulatePreemption, one per flavor-resource a flavor walk may touch)
// are independent: k_search_cells evaluates them all up front, one warp each, and memoises (cell, quantity) ->
// (mode, borrow height); k_nominate_walk then replays findFlavorForPodSets sequentially against the memo and
// runs GetTargets for the chosen assignment.
#pragma once

#include <climits>

#include "kb_device.cuh"

enum { PV_NEVER = 0, PV_WITHIN_CQ = 1, PV_HIER_RECLAIM = 2, PV_RECLAIM_NO_BORROW = 3, PV_RECLAIM_WHILE_BORROW = 4 };
enum { TC_USE = 1, TC_NEED = 2 };

__device__ __forceinline__ int variant_reason(int v) {  // PreemptionReason hierarchical_preemption.go:49-61
  switch (v) {
    case PV_WITHIN_CQ: return KB_REASON_IN_CLUSTER_QUEUE;
    case PV_HIER_RECLAIM: return KB_REASON_IN_COHORT_RECLAMATION;
    case PV_RECLAIM_WHILE_BORROW: return KB_REASON_IN_COHORT_RECLAIM_WHILE_BORROWING;
    case PV_RECLAIM_NO_BORROW: return KB_REASON_IN_COHORT_RECLAMATION;
  }
  return 0;
}

// ---------------------------------------------------------------------------
// Per-cycle search tables.
// ---------------------------------------------------------------------------
// Nodes of one root in "slot-node" numbering: g = slot_base[slot] + h, h = local index (depth-ascending order inside
// a cohort tree, 0 for a cohort-less ClusterQueue).  Column tables are stored transposed per slot:
//   index(slot, fr, h) = slot_base[slot] * FR + fr * nn + h
// so one column of one root is contiguous.
__global__ void k_columns(DevSnap D) {
  const int FR = D.FR;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)D.N * FR) return;
  int g = (int)(i / FR), fr = (int)(i % FR);  // thread -> (slot-node, fr): reads of the [node][FR] tables are coalesced
  int node = D.sn_node[g];
  int slot = D.root_slot[node];
  int base = D.slot_base[slot];
  int nn = D.slot_base[slot + 1] - base;
  int h = g - base;
  size_t c = (size_t)node * FR + fr;
  i64 sub = D.subtree[c], u = D.usage[c];
  ColStat st;
  st.sub = sub; st.lq = local_quota(sub, D.llimit[c]); st.bl = D.blimit[c];
  int p = D.parent[node];
  st.parent = p < 0 ? -1 : D.local_idx[p];
  st.depth = (int16_t)D.depth[node]; st.height = (int16_t)D.height[node];
  size_t o = (size_t)base * FR + (size_t)fr * nn + h;
  D.colU[o] = u;
  D.colS[o] = st;
  // bit d (d < depth): the ancestor at depth d is above nominal in fr; bit depth: the node itself
  // (IsWithinNominalInResources resource_node.go:248-255 is the negation, per flavor-resource)
  uint32_t m = u > sub ? 1u << D.depth[node] : 0u;
  for (int t = p; t >= 0; t = D.parent[t]) {
    size_t tc = (size_t)t * FR + fr;
    if (D.usage[tc] > D.subtree[tc]) m |= 1u << D.depth[t];
  }
  D.ovm[o] = m;
}

// fr-lists: bucket (slot, fr) = the admitted workloads of the root that use fr, in adm_sorted (rank) order.
__global__ void k_frl_count(DevSnap D) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= D.AU) return;
  // owner of usage cell k: binary search in adm_use_start
  int lo = 0, hi = D.A;
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (D.adm_use_start[mid] <= k) lo = mid; else hi = mid; }
  int slot = D.root_slot[D.adm_cq[lo]];
  atomicAdd(&D.frl_count[(size_t)slot * D.FR + D.adm_use_fr[k]], 1);
}
// exclusive scan of n int32 counters (single CTA; n is nRoots*FR or nRoots)
__global__ void __launch_bounds__(1024) k_scan_i32(const int32_t *in, int32_t *out, int n) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < n ? in[i] : 0;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (w == 0) {
      int s = warp_sums[lane];
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
      warp_sums[lane] = s;
    }
    __syncthreads();
    int excl = carry + (w ? warp_sums[w - 1] : 0) + x - v;
    if (i < n) out[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}
// one warp per bucket: ordered (ballot) compaction of the root's ranked list
__global__ void __launch_bounds__(128) k_frl_fill(DevSnap D) {
  const int FR = D.FR;
  int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (b >= D.nRoots * FR) return;
  int n = D.frl_start[b + 1] - D.frl_start[b];
  if (n == 0) return;
  int slot = b / FR, fr = b % FR;
  int base = D.slot_base[slot], nn = D.slot_base[slot + 1] - base;
  const int32_t *list = D.adm_sorted + D.root_adm_start[slot];
  int len = D.root_adm_start[slot + 1] - D.root_adm_start[slot];
  FrRec *out = D.frl + D.frl_start[b];
  int written = 0;
  for (int i0 = 0; i0 < len && written < n; i0 += 32) {
    int i = i0 + lane;
    int a = -1; i64 qty = 0; bool has = false;
    if (i < len) {
      a = list[i];
      for (int k = D.adm_use_start[a]; k < D.adm_use_start[a + 1]; k++)
        if (D.adm_use_fr[k] == fr) { has = true; qty = D.adm_use_qty[k]; break; }
    }
    unsigned m = __ballot_sync(0xffffffffu, has);
    if (has) {
      int cq = D.adm_cq[a];
      int h = D.local_idx[cq];
      FrRec r;
      r.adm = a; r.hcq = h; r.prio = D.adm_priority[a];
      r.info = (D.adm_evicted[a] ? 1u : 0u) | ((uint32_t)D.depth[cq] << 1) | (D.ovm[(size_t)base * FR + (size_t)fr * nn + h] << 8);
      r.qty = qty; r.tin = D.nd_tin[base + h]; r.cq = cq;
      out[written + __popc(m & ((1u << lane) - 1))] = r;
    }
    written += __popc(m);
  }
}
// The root's ranked list as packed records (multi-cell GetTargets searches stream these instead of gathering):
// same 32 B layout as FrRec with the quantity replaced by the set of flavor-resources the workload uses (bit fr,
// exact when F*R <= 64, all ones otherwise -> the CSR cells are consulted).
__global__ void k_root_recs(DevSnap D) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.A) return;
  int a = D.adm_sorted[i];
  int cq = D.adm_cq[a];
  int slot = D.root_slot[cq];
  int h = D.local_idx[cq];
  u64 mask = 0;
  if (D.FR <= 64) { for (int k = D.adm_use_start[a]; k < D.adm_use_start[a + 1]; k++) mask |= 1ull << D.adm_use_fr[k]; }
  else mask = ~0ull;
  FrRec r;
  r.adm = a; r.hcq = h; r.prio = D.adm_priority[a];
  r.info = (D.adm_evicted[a] ? 1u : 0u) | ((uint32_t)D.depth[cq] << 1);
  r.qty = (i64)mask; r.tin = D.nd_tin[D.slot_base[slot] + h]; r.cq = cq;
  D.rrec[i] = r;
}

// ---------------------------------------------------------------------------
// Column arithmetic on a private copy of one flavor-resource column (resource_node.go).
// ---------------------------------------------------------------------------
__device__ __forceinline__ ColStat ld_stat(const ColStat *s, int h) {
  const int4 *p = reinterpret_cast<const int4 *>(s + h);  // generic loads: the column may sit in shared memory
  int4 a = p[0], b = p[1];
  ColStat st;
  st.sub = ((i64)(unsigned)a.y << 32) | (unsigned)a.x; st.lq = ((i64)(unsigned)a.w << 32) | (unsigned)a.z;
  st.bl = ((i64)(unsigned)b.y << 32) | (unsigned)b.x; st.parent = b.z;
  st.depth = (int16_t)(b.w & 0xffff); st.height = (int16_t)((unsigned)b.w >> 16);
  return st;
}
__device__ inline void col_remove(i64 *c, const ColStat *s, int h, i64 val) {  // removeUsage :149-158
  while (true) {
    ColStat st = ld_stat(s, h);
    i64 u = c[h], stored = u - st.lq;
    c[h] = u - val;
    if (stored <= 0 || st.parent < 0) break;
    val = imin(val, stored); h = st.parent;
  }
}
__device__ inline void col_add(i64 *c, const ColStat *s, int h, i64 val) {  // addUsage :137-145
  while (true) {
    ColStat st = ld_stat(s, h);
    i64 u = c[h], la = imax(0, st.lq - u);
    c[h] = u + val;
    if (st.parent < 0 || !(val > la)) break;
    val -= la; h = st.parent;
  }
}
// available :104-118 along the staged path (path[0] = ClusterQueue ... path[plen-1] = root), clamped like
// ClusterQueueSnapshot.Available
__device__ inline i64 col_avail(const i64 *c, const ColStat *s, const int *path, int plen) {
  int rt = path[plen - 1];
  i64 a = ld_stat(s, rt).sub - c[rt];
  for (int k = plen - 2; k >= 0; k--) {
    int nd = path[k];
    ColStat st = ld_stat(s, nd);
    i64 u = c[nd], pa = a;
    if (st.bl != KB_NO_LIMIT) pa = imin((st.sub - st.lq) - imax(0, u - st.lq) + st.bl, pa);
    a = imax(0, st.lq - u) + pa;
  }
  return imax(0, a);
}
// FindHeightOfLowestSubtreeThatFits hierarchical_preemption.go:214-227 on the private column
__device__ inline int col_find_height(const i64 *c, const ColStat *s, int h, i64 val) {
  ColStat st = ld_stat(s, h);
  if (!(c[h] + val > st.sub) || st.parent < 0) return 0;
  i64 remaining = val - imax(0, st.lq - c[h]);
  int t = st.parent, last = t;
  while (t >= 0) {
    st = ld_stat(s, t);
    if (!(c[t] + remaining > st.sub)) return st.height;
    remaining -= imax(0, st.lq - c[t]);
    last = t; t = st.parent;
  }
  return ld_stat(s, last).height;
}
// quantity of admitted workload a in flavor-resource fr (Info.FlavorResourceUsage is a map: one cell per fr)
__device__ __forceinline__ i64 adm_qty(const DevSnap &D, int a, int fr) {
  for (int k = D.adm_use_start[a]; k < D.adm_use_start[a + 1]; k++) if (D.adm_use_fr[k] == fr) return D.adm_use_qty[k];
  return 0;
}

// ---------------------------------------------------------------------------
// Warp search context (shared memory, one per warp).  KMAX = tracked columns it can describe.
// ---------------------------------------------------------------------------
template <int KMAX>
struct WCtx {
  int slot, nn, node_base, plen, K, n_need, cq, hcq, prio;
  i64 ts;
  size_t cbase;                 // slot_base[slot] * FR
  int path[KB_MAX_DEPTH + 1];   // local handles, path[0] = the preemptor's ClusterQueue
  int ptin[KB_MAX_DEPTH + 1], ptout[KB_MAX_DEPTH + 1];
  uint8_t adv[KB_MAX_DEPTH + 1];  // hierarchical advantage when collecting at path level k
  uint16_t tfr[KMAX];           // tracked column -> flavor-resource
  uint8_t tflag[KMAX];          // TC_USE | TC_NEED
  i64 tq[KMAX];                 // workloadUsage.Quota of the column (0 when not TC_USE)
  uint32_t need_bits[(KMAX > 1) ? 32 : 1];  // flavor-resources needing preemption, as a bit set (multi-cell searches)
};

// per-warp scratch handed to a search
struct WScratch {
  i64 *col;            // [K][nn] private columns (shared memory when it fits, else global)
  uint8_t *codes;      // [>= list length] (segment, variant) code per candidate
  int32_t *tgt;        // [>= list length] targets of the current search
  uint8_t *tgt_reason;
  i64 *tgtq = nullptr; // [>= bucket length] quantity of every target in the searched cell (single-column searches)
  // single-column searches of a CTA bound to one (root, flavor-resource) bucket: that column's static cell records
  // and cycle-start usage staged once in shared memory (nullptr: read the global tables)
  const ColStat *stat0 = nullptr;
  const i64 *base0 = nullptr;
};

__device__ __forceinline__ bool ws_satisfies_policy(const DevSnap &D, int pre_prio, i64 pre_ts, int a, int cp, int policy) {  // preemption_policy.go:30-48
  bool lower = pre_prio > cp;
  if (policy == KB_POLICY_LOWER_PRIORITY) return lower;
  if (policy == KB_POLICY_LOWER_OR_NEWER_EQUAL_PRIORITY) {
    if (lower) return true;
    if (pre_prio != cp) return false;
    i64 ats = D.adm_ts[a];
    bool newer = pre_ts < ats;
    if (newer && (D.flags & KB_F_TS_PREEMPTION_BUFFER)) newer = ats - pre_ts > 300ll * 1000000000ll;  // timestampPreemptionBuffer :28
    return newer;
  }
  return policy == KB_POLICY_ANY;
}

// classicalPreemptions preemption.go:238-293 for the context in *w (cq, prio, ts, tracked columns filled in; every lane
// passes the quantities of the columns it owns: column j = lane + 32 * s -> myq[s]).  All 32 lanes call it together.
// Returns the number of targets; they are left in S.tgt / S.tgt_reason.  The private columns are NOT restored (every
// search starts by loading them).
template <int KMAX, int KPL>
__device__ inline int ws_classical(const DevSnap &D, WCtx<KMAX> *w, const WScratch &S, const i64 (&myq)[KPL]) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int FR = D.FR;
  const int cq = w->cq, K = w->K;
  const int slot = D.root_slot[cq];
  const int nbase = D.slot_base[slot], nn = D.slot_base[slot + 1] - nbase;
  const size_t cbase = (size_t)nbase * FR;
  const int hcq = D.local_idx[cq];
  const bool has_parent = D.parent[cq] >= 0;
  const int pol_within = D.cq_within_cq[cq], pol_reclaim = D.cq_reclaim_within[cq];
  const bool cohort_cands = has_parent && pol_reclaim != KB_POLICY_NEVER;
  const bool own_cands = pol_within != KB_POLICY_NEVER;
  const int prio = w->prio; const i64 ts = w->ts;
  // ---- lane-owned columns
  int myfr[KPL]; int myflag[KPL]; i64 *mycol[KPL]; const ColStat *mys[KPL];
#pragma unroll
  for (int s = 0; s < KPL; s++) {
    int j = lane + 32 * s;
    bool on = j < K;
    myfr[s] = on ? w->tfr[j] : 0; myflag[s] = on ? w->tflag[j] : 0;
    mycol[s] = S.col + (size_t)(on ? j : 0) * nn;
    mys[s] = (S.stat0 && K == 1) ? S.stat0 : D.colS + cbase + (size_t)myfr[s] * nn;
  }
  // ---- 1. preemptor path (lane 0), Euler intervals of the path nodes
  if (lane == 0) {
    const ColStat *s0 = (S.stat0 && K == 1) ? S.stat0 : D.colS + cbase + (size_t)w->tfr[0] * nn;
    int pl = 0;
    for (int t = hcq; t >= 0; t = ld_stat(s0, t).parent) {
      w->path[pl] = t; w->ptin[pl] = D.nd_tin[nbase + t]; w->ptout[pl] = D.nd_tout[nbase + t]; pl++;
    }
    w->plen = pl; w->hcq = hcq; w->nn = nn; w->slot = slot; w->node_base = nbase; w->cbase = cbase;
  }
  // ---- 2. private columns <- cycle-start usage
  auto load_columns = [&]() {
    for (int j = 0; j < K; j++) {
      const i64 *src = (S.base0 && K == 1) ? S.base0 : D.colU + cbase + (size_t)w->tfr[j] * nn;
      i64 *dst = S.col + (size_t)j * nn;
      for (int h = lane; h < nn; h += 32) dst[h] = src[h];
    }
    __syncwarp();
  };
  __syncwarp();
  long long t0 = clock64();
  load_columns();
  long long t1 = clock64();
  const int plen = w->plen;
  const int *path = w->path;
  // ---- 3. hierarchical advantage per path level (collectCandidatesForHierarchicalReclaim :151-177,
  //         QuantitiesFitInQuota resource_node.go:234-244): lane j walks its column
  {
    i64 rem[KPL];
#pragma unroll
    for (int s = 0; s < KPL; s++) rem[s] = myq[s];
    bool adv = false;
    for (int k = 0; k < plen; k++) {
      int nd = path[k];
      bool fits = true;
#pragma unroll
      for (int s = 0; s < KPL; s++) {
        if (!(myflag[s] & TC_USE)) continue;
        ColStat st = ld_stat(mys[s], nd);
        i64 u = mycol[s][nd];
        if (u + rem[s] > st.sub) fits = false;
        rem[s] = imax(0, rem[s] - imax(0, st.lq - u));
      }
      fits = __all_sync(FULL, fits);
      if (lane == 0) w->adv[k] = (uint8_t)adv;  // advantage in force while collecting at level k (k >= 1)
      adv = adv || fits;
    }
  }
  __syncwarp();
  // ---- 4. classify every candidate of the list: code = segment * 8 + variant, 0xff = not a candidate
  const bool single = w->n_need == 1;
  int list_off, list_len;
  if (single) {
    int nfr = 0;
    for (int j = 0; j < K; j++) if (w->tflag[j] & TC_NEED) nfr = w->tfr[j];
    int b = slot * FR + nfr;
    list_off = D.frl_start[b]; list_len = D.frl_start[b + 1] - list_off;
  } else {
    list_off = D.root_adm_start[slot]; list_len = D.root_adm_start[slot + 1] - list_off;
  }
  const bool forbidden = D.cq_borrow_within[cq] == KB_POLICY_NEVER;  // IsBorrowingWithinCohortForbidden :72-78
  const bool has_thr = D.cq_has_bwc_threshold[cq]; const int thr = D.cq_bwc_threshold[cq];
  unsigned present = 0;  // bit seg: the segment has a candidate
  // Candidates are streamed as 32 B records in rank order (evicted first, then priority ascending): two batches of
  // 32 in flight per step.  No candidate above `pmax` can satisfy either preemption policy
  // (preemption_policy.go:30-48), so the stream ends at the first non-evicted batch that starts above it.
  long long pmax = LLONG_MIN;
  {
    auto cap = [&](int pol) -> long long {
      return pol == KB_POLICY_LOWER_PRIORITY ? (long long)prio - 1 : pol == KB_POLICY_LOWER_OR_NEWER_EQUAL_PRIORITY ? (long long)prio
             : pol == KB_POLICY_ANY ? LLONG_MAX : LLONG_MIN;
    };
    if (own_cands) pmax = cap(pol_within);
    if (cohort_cands) { long long c2 = cap(pol_reclaim); if (c2 > pmax) pmax = c2; }
  }
  const FrRec *recs = single ? D.frl + list_off : D.rrec + list_off;
  const bool mask_exact = FR <= 64;
  u64 need64 = 0;
  if (!single) need64 = (u64)w->need_bits[0] | ((u64)w->need_bits[1] << 32);
  auto classify = [&](int4 r0, int4 r1) -> int {
    int a = r0.x, h = r0.y, cp = r0.z; uint32_t info = (uint32_t)r0.w;
    bool evicted = info & 1; int depth = (info >> 1) & 31; int tin = r1.z;
    uint32_t ovmask = info >> 8;
    bool same = h == hcq;
    if (!single) {
      bool uses;  // WorkloadUsesResources candidate_generator.go:52-61
      u64 um = ((u64)(unsigned)r1.y << 32) | (unsigned)r1.x;
      if (mask_exact) uses = (um & need64) != 0;
      else {
        uses = false;
        for (int k = D.adm_use_start[a]; k < D.adm_use_start[a + 1]; k++) { int fr = D.adm_use_fr[k]; if ((w->need_bits[fr >> 5] >> (fr & 31)) & 1) uses = true; }
      }
      if (!uses) return 0xff;
    }
    if (!(same ? own_cands : cohort_cands)) return 0xff;
    if (!ws_satisfies_policy(D, prio, ts, a, cp, same ? pol_within : pol_reclaim)) return 0xff;  // classifyPreemptionVariant :82-114
    int variant, cls;
    if (same) { variant = PV_WITHIN_CQ; cls = 2; }
    else {
      if (!single) { ovmask = 0; for (int j = 0; j < K; j++) if (w->tflag[j] & TC_NEED) ovmask |= D.ovm[cbase + (size_t)w->tfr[j] * nn + h]; }
      if (!((ovmask >> depth) & 1)) return 0xff;  // the ClusterQueue is within nominal in every cell needing preemption
      int lvl = plen - 1;                         // lowest common ancestor with the preemptor's path
      for (int k = 1; k < plen; k++) if (tin >= w->ptin[k] && tin < w->ptout[k]) { lvl = k; break; }
      int dt = plen - 1 - lvl;                    // its depth; the cohorts strictly between must be above nominal too
      uint32_t chain = ((1u << depth) - 1u) & ~((1u << (dt + 1)) - 1u);
      if ((ovmask & chain) != chain) return 0xff;
      bool hier = w->adv[lvl];
      cls = hier ? 0 : 1;
      if (hier) variant = PV_HIER_RECLAIM;
      else if (forbidden) variant = PV_RECLAIM_NO_BORROW;
      else {
        bool above;  // isAboveBorrowingThreshold :116-124
        if (cp >= prio) above = true;
        else if (!has_thr) above = false;
        else above = cp > thr;
        variant = above ? PV_RECLAIM_NO_BORROW : PV_RECLAIM_WHILE_BORROW;
      }
    }
    int seg = (evicted ? 0 : 3) + cls;
    present |= 1u << seg;
    return seg * 8 + variant;
  };
  int scan_len = list_len;
  for (int i0 = 0; i0 < list_len; i0 += 64) {
    const int ia = i0 + lane, ib = i0 + 32 + lane;
    int4 a0 = make_int4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
    if (ia < list_len) { const int4 *rp = reinterpret_cast<const int4 *>(recs + ia); a0 = __ldg(rp); a1 = __ldg(rp + 1); }
    if (ib < list_len) { const int4 *rp = reinterpret_cast<const int4 *>(recs + ib); b0 = __ldg(rp); b1 = __ldg(rp + 1); }
    // first record of the step: not evicted and already above every admissible priority -> nothing further qualifies
    int first_prio = __shfl_sync(FULL, a0.z, 0); unsigned first_info = (unsigned)__shfl_sync(FULL, a0.w, 0);
    if (!(first_info & 1) && (long long)first_prio > pmax) { scan_len = i0; break; }
    if (ia < list_len) S.codes[ia] = (uint8_t)classify(a0, a1);
    if (ib < list_len) S.codes[ib] = (uint8_t)classify(b0, b1);
  }
  list_len = scan_len;
  present = __reduce_or_sync(FULL, present);
  __syncwarp();
  long long t2 = clock64();
  if (lane == 0) {
    atomicAdd(&D.sstat[0], 1ull); atomicAdd(&D.sstat[1], (u64)list_len);
    atomicAdd(&D.sstat[5], (u64)(t1 - t0)); atomicAdd(&D.sstat[6], (u64)(t2 - t1));
    if (K > 1) atomicAdd(&D.sstat[4], (u64)list_len);
  }
  if (present == 0) return 0;
  int n_visit = 0, n_removed = 0;
  // ---- 5. greedy remove / fill back (preemption.go:265-293)
  const bool no_hier = !(present & 0x09), no_other = !(present & 0x1b);
  bool under_nominal = true;  // queueUnderNominalInResourcesNeedingPreemption :577-584
#pragma unroll
  for (int s = 0; s < KPL; s++)
    if (myflag[s] & TC_NEED) { if (mycol[s][hcq] >= ld_stat(mys[s], hcq).sub) under_nominal = false; }
  under_nominal = __all_sync(FULL, under_nominal);
  bool opts[2]; int nopts;
  if (no_other || (forbidden && !under_nominal)) { opts[0] = true; nopts = 1; }   // :266-267
  else if (forbidden && no_hier) { opts[0] = false; opts[1] = true; nopts = 2; }  // :268-269
  else { opts[0] = true; opts[1] = false; nopts = 2; }                            // :270-271
  auto within_nominal = [&](int h) {  // IsWithinNominalInResources on the private columns, warp-uniform
    bool over = false;
#pragma unroll
    for (int s = 0; s < KPL; s++) if (myflag[s] & TC_NEED) over = over || mycol[s][h] > ld_stat(mys[s], h).sub;
    return !__any_sync(FULL, over);
  };
  auto fits = [&](bool allow_borrowing) {  // workloadFits :550-561
    bool ok = true;
#pragma unroll
    for (int s = 0; s < KPL; s++) {
      if (!(myflag[s] & TC_USE)) continue;
      if (!allow_borrowing && mycol[s][hcq] + myq[s] > ld_stat(mys[s], hcq).sub) ok = false;
      else if (myq[s] > col_avail(mycol[s], mys[s], path, plen)) ok = false;
    }
    return (bool)__all_sync(FULL, ok);
  };
  // Snapshot.RemoveWorkload / AddWorkload on the tracked columns; q0 >= 0: the quantity in column 0 is already known
  // (single-column searches carry it in the bucket record)
  auto apply_adm = [&](int a, int h, bool remove, i64 q0 = -1) {
#pragma unroll
    for (int s = 0; s < KPL; s++) {
      if (!myflag[s]) continue;
      i64 q = (q0 >= 0 && s == 0 && K == 1) ? q0 : adm_qty(D, a, myfr[s]);
      if (q == 0) continue;
      if (remove) col_remove(mycol[s], mys[s], h, q); else col_add(mycol[s], mys[s], h, q);
    }
  };
  const ColStat *s0 = (S.stat0 && K == 1) ? S.stat0 : D.colS + cbase + (size_t)w->tfr[0] * nn;  // parent / depth of any node (identical in every column)
  int nt = 0; bool found = false;
  for (int oi = 0; oi < nopts && !found; oi++) {
    const bool borrow = opts[oi];
    if (oi > 0) load_columns();  // restoreSnapshot :310-314
    nt = 0;
    for (int seg = 0; seg < 6 && !found; seg++) {
      if (!((present >> seg) & 1)) continue;
      int nxt = lane < list_len ? S.codes[lane] : 0xff;
      for (int i0 = 0; i0 < list_len && !found; i0 += 32) {
        int cur = nxt;
        nxt = (i0 + 32 + lane < list_len) ? S.codes[i0 + 32 + lane] : 0xff;
        const bool mine = cur != 0xff && (cur >> 3) == seg;
        unsigned m = __ballot_sync(FULL, mine);
        if (!m) continue;
        // the records of every candidate of this step in one coalesced load; the ordered walk below reads them by shuffle
        int my_a = 0, my_h = 0; i64 my_q = -1;
        if (mine) {
          const int4 *rp = reinterpret_cast<const int4 *>(recs + i0 + lane);
          int4 r0 = __ldg(rp); my_a = r0.x; my_h = r0.y;
          if (single && K == 1) { int4 r1 = __ldg(rp + 1); my_q = ((i64)(unsigned)r1.y << 32) | (unsigned)r1.x; }
        }
        while (m) {
          int src = __ffs(m) - 1; m &= m - 1;
          int v = __shfl_sync(FULL, cur, src) & 7;
          n_visit++;
          const int a = __shfl_sync(FULL, my_a, src), h = __shfl_sync(FULL, my_h, src);
          const i64 q0 = __shfl_sync(FULL, my_q, src);
          if (h != hcq) {  // candidateIsValid candidate_generator.go:140-162
            if (borrow && v == PV_RECLAIM_NO_BORROW) continue;
            if (within_nominal(h)) continue;
            bool valid = true;
            for (int t = ld_stat(s0, h).parent; t >= 0; ) {
              ColStat st = ld_stat(s0, t);
              int kp = plen - 1 - st.depth;
              if (kp >= 0 && path[kp] == t) break;  // reached the subtree root that collected the candidate
              if (within_nominal(t)) { valid = false; break; }
              t = st.parent;
            }
            if (!valid) continue;
          }
          apply_adm(a, h, true, q0);
          n_removed++;
          if (lane == 0) { S.tgt[nt] = a; S.tgt_reason[nt] = (uint8_t)variant_reason(v); if (S.tgtq && K == 1) S.tgtq[nt] = q0; }
          nt++;
          if (fits(borrow)) {
            __syncwarp();
            for (int k = nt - 2; k >= 0; k--) {  // fillBackWorkloads :295-308
              int b = S.tgt[k];
              i64 qb = (S.tgtq && K == 1) ? S.tgtq[k] : -1;
              int hb = D.local_idx[D.adm_cq[b]];
              apply_adm(b, hb, false, qb);
              n_removed++;
              if (fits(borrow)) {
                if (lane == 0) { S.tgt[k] = S.tgt[nt - 1]; S.tgt_reason[k] = S.tgt_reason[nt - 1]; if (S.tgtq && K == 1) S.tgtq[k] = S.tgtq[nt - 1]; }
                nt--;
                __syncwarp();
              } else apply_adm(b, hb, true, qb);
            }
            found = true;
            break;
          }
        }
      }
    }
  }
  __syncwarp();
  if (lane == 0) {
    atomicAdd(&D.sstat[2], (u64)n_visit); atomicAdd(&D.sstat[3], (u64)n_removed);
    atomicAdd(&D.sstat[7], (u64)(clock64() - t2));
  }
  return found ? nt : 0;
}

// SimulatePreemption preemption_oracle.go:41-71 for one flavor-resource: all lanes call it together.
template <int KMAX>
__device__ inline int ws_simulate(const DevSnap &D, WCtx<KMAX> *w, const WScratch &S, int wl, int cq, int fr, i64 val, int *borrow_after) {
  const int lane = threadIdx.x & 31;
  bool may_reclaim;
  *borrow_after = find_height(D, D.usage, cq, fr, val, &may_reclaim);  // no candidates: height on the untouched snapshot (:53-56)
  __syncwarp();
  if (lane == 0) {
    w->cq = cq; w->prio = D.wl_priority[wl]; w->ts = D.wl_ts[wl];
    w->K = 1; w->n_need = 1; w->tfr[0] = (uint16_t)fr; w->tflag[0] = TC_USE | TC_NEED; w->tq[0] = val;
  }
  __syncwarp();
  i64 myq[1] = {lane == 0 ? val : 0};
  int nt = ws_classical<KMAX, 1>(D, w, S, myq);
  if (nt == 0) return 1;  // PM_NOCAND
  // borrow height with the targets removed (:57-63); the column is private and reloaded by the next search
  int slot = D.root_slot[cq];
  int nbase = D.slot_base[slot], nn = D.slot_base[slot + 1] - nbase;
  const ColStat *s0 = S.stat0 ? S.stat0 : D.colS + (size_t)nbase * D.FR + (size_t)fr * nn;
  int hcq = D.local_idx[cq];
  int b = 0; bool own = false;
  if (lane == 0) {
    // the search left the column with exactly the targets removed
    b = col_find_height(S.col, s0, hcq, val);
    for (int k = 0; k < nt; k++) if (D.adm_cq[S.tgt[k]] == cq) own = true;
  }
  b = __shfl_sync(0xffffffffu, b, 0);
  own = __shfl_sync(0xffffffffu, (int)own, 0);
  *borrow_after = b;
  return own ? 2 : 3;  // PM_PREEMPT : PM_RECLAIM
}
