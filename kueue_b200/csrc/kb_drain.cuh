// kb_drain.cuh — the queue layer of kb_run_drain on the device.
//
// Reference: pkg/cache/queue/cluster_queue.go:362-425 (requeueIfNotPresent, handleInadmissibleHash), :494-510 (Pop),
// :609-685 (RequeueIfNotPresent, queueOrderingFunc); manager.go:770-794 (heads); pkg/scheduler/scheduler.go:405-418,
// 823-850 (requeueAndUpdate); pkg/workload/workload.go:161-191 (PendingFlavors, NextFlavorToTryForPodSetResource).
//
// The heaps of the reference become, per ClusterQueue, one segment of the pending workloads sorted once by
// queueOrderingFunc (priority desc, queue-order timestamp, UID) plus a cursor: inside a drain nothing re-enters a
// queue (no cluster events), so "pop the heap" is "first queued workload at or after the cursor that was not moved
// to the inadmissible set".  Every cycle: pick the heads (ordered compaction over ClusterQueues), run the unchanged
// cycle kernels on them, then apply the decisions — admissions extend the admitted tables and the ClusterQueue
// usage on the device, the others keep their LastAssignment and move per the Strict / BestEffort FIFO rules.
#pragma once

#include "kb_device.cuh"

struct DrainDev {
  // queues
  int32_t *q_order, *q_start, *cursor; uint8_t *gone;
  int32_t *flag, *pos;                       // [Q+1] ClusterQueues with work left / their entry position
  int32_t *e_assumed, *e_ncells, *e_adm_off, *e_cell_off;  // [Hcap+1]
  int32_t *counters;                          // [0] entries of the next cycle
  // per pending workload
  int32_t *wl_admit_cycle, *wl_evals; uint8_t *wl_last_decision;
  int32_t *trace_wl; uint8_t *trace_dec; long long trace_off, trace_cap;
  // mutable views of the snapshot tables
  i64 *cq_usage, *wl_last_gen; int8_t *ps_last_tried;
  int32_t *adm_cq, *adm_priority; i64 *adm_ts, *adm_qr_ts, *adm_uid; uint8_t *adm_evicted; int32_t *adm_use_start, *adm_use_fr; i64 *adm_use_qty;
  int A, AU, cycle;  // admitted workloads / usage cells before this cycle's admissions
};

// sort keys of queueOrderingFunc (cluster_queue.go:636-685): LSD passes uid, timestamp, (ClusterQueue | priority desc)
__global__ void k_drain_keys_uid(DevSnap D, u64 *keys, int32_t *vals) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= D.W) return;
  keys[w] = (u64)D.wl_uid[w] ^ 0x8000000000000000ull;
  vals[w] = w;
}
__global__ void k_drain_keys_ts(DevSnap D, const int32_t *vals, u64 *keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.W) return;
  keys[i] = (u64)D.wl_ts[vals[i]] ^ 0x8000000000000000ull;
}
__global__ void k_drain_keys_cq(DevSnap D, const int32_t *vals, u64 *keys, int32_t *q_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.W) return;
  int w = vals[i];
  int cq = D.wl_cq[w];
  keys[i] = ((u64)cq << 32) | (u64)(~((unsigned)D.wl_priority[w] ^ 0x80000000u));  // higher priority first
  atomicAdd(&q_count[cq], 1);
}
__global__ void k_drain_init(DevSnap D, DrainDev X) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < D.W) { X.gone[i] = 0; X.wl_admit_cycle[i] = -1; X.wl_evals[i] = 0; X.wl_last_decision[i] = 0xff; }
  if (i < D.Q) X.cursor[i] = 0;
}

// heads of the next cycle (queues.Heads manager.go:770-794): one per ClusterQueue with work left, in ClusterQueue order
__global__ void k_drain_flag(DevSnap D, DrainDev X) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= D.Q) return;
  X.flag[q] = X.cursor[q] < X.q_start[q + 1] - X.q_start[q];
}
__global__ void k_drain_heads(DevSnap D, DrainDev X, int32_t *heads) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= D.Q) return;
  if (X.flag[q]) heads[X.pos[q]] = X.q_order[X.q_start[q] + X.cursor[q]];
  if (q == 0) X.counters[0] = X.pos[D.Q];
}

__device__ __forceinline__ bool drain_covers_pods(const DevSnap &D, int cq) {
  if (D.pods_res < 0) return false;
  for (int g = D.cq_rg_start[cq]; g < D.cq_rg_start[cq + 1]; g++) if (D.rg_res_mask[g] & (1u << D.pods_res)) return true;
  return false;
}
__device__ __forceinline__ i64 drain_request(const DevSnap &D, int row, int r, int count, bool covers_pods) {  // ScaledTo workload.go:258-275
  if (covers_pods && r == D.pods_res) return count;
  i64 q = D.ps_req[(size_t)row * D.R + r];
  int full = D.ps_count[row];
  if (full != 0 && full != count) q = q / full * count;
  return q;
}

// decisions of the cycle -> LastAssignment, queue movement, per-entry admission bookkeeping
__global__ void k_drain_apply(DevSnap D, DrainDev X) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= D.H) return;
  const int R = D.R;
  int wl = D.heads[e], cq = D.wl_cq[wl];
  int dec = D.decision[e];
  X.wl_last_decision[wl] = (uint8_t)dec;
  X.wl_evals[wl]++;
  if (X.trace_wl && X.trace_off + e < X.trace_cap) { X.trace_wl[X.trace_off + e] = wl; X.trace_dec[X.trace_off + e] = (uint8_t)dec; }
  const bool ok = dec == KB_DEC_ASSUMED;
  const int ps0 = D.wl_ps_start[wl], ps1 = D.wl_ps_start[wl + 1];
  int ncell = 0;
  bool pending_flavors = false;
  if (ok) {
    X.wl_admit_cycle[wl] = X.cycle;
    for (int row = ps0; row < ps1; row++)  // distinct flavor-resource cells of Assignment.Usage
      for (int r = 0; r < R; r++) {
        int f = D.ps_flavor[(size_t)row * R + r];
        if (f < 0) continue;
        bool first = true;  // same cell = same resource and same flavor in an earlier podset
        for (int prow = ps0; prow < row; prow++) if (D.ps_flavor[(size_t)prow * R + r] == f) { first = false; break; }
        if (first) ncell++;
      }
  } else if (dec == KB_DEC_PREEMPTING) {
    X.wl_last_gen[wl] = -1;  // "the next attempt should try all the flavors" scheduler.go:345
  } else {
    X.wl_last_gen[wl] = D.cq_generation[cq];  // e.LastAssignment = &e.assignment.LastState scheduler.go:494
    for (int i = ps0 * R; i < ps1 * R; i++) { int8_t t = D.ps_tried[i]; X.ps_last_tried[i] = t; if (t != -1) pending_flavors = true; }  // PendingFlavors workload.go:163-176
  }
  X.e_assumed[e] = ok; X.e_ncells[e] = ncell;
  // ---- queue movement (cluster_queue.go:362-425,609-633)
  const bool strict = D.cq_strategy[cq] == KB_QUEUE_STRICT_FIFO;
  const bool inadmissible = (dec == KB_DEC_NOFIT || dec == KB_DEC_PREEMPT_NO_TARGETS) && !pending_flavors;
  const int qs = X.q_start[cq], qn = X.q_start[cq + 1] - qs;
  int cur = X.cursor[cq];
  if (dec == KB_DEC_NOFIT && !strict && D.wl_sched_hash) {  // handleInadmissibleHash :408-425
    i64 hsh = D.wl_sched_hash[wl];
    if (hsh != 0)
      for (int p = cur + 1; p < qn; p++) { int w2 = X.q_order[qs + p]; if (D.wl_sched_hash[w2] == hsh) X.gone[w2] = 1; }
  }
  if (ok || (inadmissible && !strict)) cur++;
  while (cur < qn && X.gone[X.q_order[qs + cur]]) cur++;
  X.cursor[cq] = cur;
}

// admitted entries join the admitted tables (cache.AssumeWorkload) and the ClusterQueue usage
__global__ void k_drain_admit(DevSnap D, DrainDev X, i64 qr_ts) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= D.H || !X.e_assumed[e]) return;
  const int R = D.R;
  int wl = D.heads[e], cq = D.wl_cq[wl];
  int a = X.A + X.e_adm_off[e];
  int c0 = X.AU + X.e_cell_off[e];
  X.adm_cq[a] = cq; X.adm_priority[a] = D.wl_priority[wl]; X.adm_ts[a] = D.wl_ts[wl]; X.adm_qr_ts[a] = qr_ts;
  X.adm_uid[a] = D.wl_uid[wl]; X.adm_evicted[a] = 0;
  X.adm_use_start[a] = c0;
  const bool covers_pods = drain_covers_pods(D, cq);
  const int ps0 = D.wl_ps_start[wl], ps1 = D.wl_ps_start[wl + 1];
  int n = 0;
  for (int row = ps0; row < ps1; row++)
    for (int r = 0; r < R; r++) {
      int f = D.ps_flavor[(size_t)row * R + r];
      if (f < 0) continue;
      int fr = f * R + r;
      i64 q = drain_request(D, row, r, D.ps_count_out[row], covers_pods);
      int j = 0;
      while (j < n && X.adm_use_fr[c0 + j] != fr) j++;
      if (j == n) { X.adm_use_fr[c0 + n] = fr; X.adm_use_qty[c0 + n] = 0; n++; }
      X.adm_use_qty[c0 + j] += q;
    }
  X.adm_use_start[a + 1] = c0 + n;
  for (int j = 0; j < n; j++) X.cq_usage[(size_t)cq * D.FR + X.adm_use_fr[c0 + j]] += X.adm_use_qty[c0 + j];  // one entry per ClusterQueue
}
