// kb_rank.cuh — device-side ranking of the admitted workloads (preemption candidates).
//
// CandidatesOrdering (pkg/scheduler/preemption/common/ordering.go:41-100) minus its preemptor-dependent key
// ("other ClusterQueues first", applied per search as a segment): evicted first, lower priority first, more
// recently reserved first (QuotaReserved transition time, now() when unset, :93-100), UID.  The order is needed
//   - per root cohort        -> adm_sorted / adm_rank (candidate buckets of the target searches, kb_search.cuh)
//   - per ClusterQueue       -> cq_adm (fair-sharing target queues, candidates_possible)
// It is a stable LSD sequence of radix sorts over (key, workload index) pairs: UID, reservation time, then
// (root | evicted | priority); one more stable pass by ClusterQueue yields the per-queue lists.  The radix sort
// itself is cub::DeviceRadixSort (plumbing, not the path's arithmetic).
#pragma once

#include <cub/device/device_radix_sort.cuh>

#include "kb_device.cuh"

__global__ void k_rank_keys_uid(DevSnap D, u64 *keys, int32_t *vals) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= D.A) return;
  keys[a] = (u64)D.adm_uid[a] ^ 0x8000000000000000ull;  // signed ascending
  vals[a] = a;
}
__global__ void k_rank_keys_qr(DevSnap D, const int32_t *vals, u64 *keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.A) return;
  int a = vals[i];
  i64 t = D.adm_qr_ts[a] == INT64_MIN ? D.now_ns : D.adm_qr_ts[a];
  keys[i] = ~((u64)t ^ 0x8000000000000000ull);  // more recent first
}
__global__ void k_rank_keys_root(DevSnap D, const int32_t *vals, u64 *keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.A) return;
  int a = vals[i];
  u64 slot = (u64)D.root_slot[D.adm_cq[a]];
  u64 prio = (u64)((unsigned)D.adm_priority[a] ^ 0x80000000u);  // lower priority first
  keys[i] = (slot << 33) | ((D.adm_evicted[a] ? 0ull : 1ull) << 32) | prio;
  atomicAdd(&D.root_adm_count[slot], 1);
}
__global__ void k_rank_keys_cq(DevSnap D, const int32_t *sorted, u64 *keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.A) return;
  int cq = D.adm_cq[sorted[i]];
  keys[i] = (u64)cq;
  atomicAdd(&D.cq_adm_count[cq], 1);
}
// position inside the root's segment
__global__ void k_rank_positions(DevSnap D) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.A) return;
  int a = D.adm_sorted[i];
  D.adm_rank[a] = i - D.root_adm_start[D.root_slot[D.adm_cq[a]]];
}
