// kueue_b200_host.hpp — C++ host side above the C-ABI (include/kueue_b200.h).
//
// The reference is Go and its toolchain is absent from this image, so the host layer a maintainer
// would write in Go is mirrored here in C++: the object model the scheduling cycle
// consumes, its flattening into kb_snapshot, and Scheduler::schedule() returning entries the way
// pkg/scheduler/scheduler.go:218-427 leaves them.  Names follow the reference:
//
//   kb::ResourceQuota / FlavorQuotas / ResourceGroup   pkg/cache/scheduler/resource.go:31-50
//   kb::ClusterQueue / Cohort                          clusterqueue_snapshot.go:37-65, cohort_snapshot.go:26-38
//   kb::PodSet / WorkloadInfo                          pkg/workload/workload.go:193-236
//   kb::Snapshot                                       pkg/cache/scheduler/snapshot.go:37-47
//   kb::Entry / PodSetAssignment / Target              scheduler.go:431-457, flavorassigner.go:262-273, preemption.go:111-115
//   kb::Scheduler::schedule                            scheduler.go:218 (decision part only; the caller replays
//                                                      admit / IssuePreemptions / requeue from the entries)
//
// Header-only, C++17, no CUDA types: link against libkueue_b200.so.  There is no CPU fallback: a non-zero status
// from the library is thrown as kb::Error and the caller runs the stock cycle.
#pragma once

#include <algorithm>
#include <cstdint>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "kueue_b200.h"

namespace kb {

struct Error : std::runtime_error {
  int32_t status;
  Error(int32_t st, const std::string &msg) : std::runtime_error(msg), status(st) {}
};

struct ResourceQuota {  // resource.go:46-50 (int64 units of resources.ResourceValue, requests.go:104-109)
  std::string resource;
  int64_t nominal = 0;
  std::optional<int64_t> borrowingLimit, lendingLimit;
};
struct FlavorQuotas {
  std::string flavor;
  std::vector<ResourceQuota> resources;
};
using ResourceGroup = std::vector<FlavorQuotas>;  // ordered flavors sharing one set of covered resources

struct ClusterQueue {
  std::string name, cohort;  // cohort == "" : no parent
  std::vector<ResourceGroup> resourceGroups;
  uint8_t withinClusterQueue = KB_POLICY_NEVER, reclaimWithinCohort = KB_POLICY_NEVER, borrowWithinCohort = KB_POLICY_NEVER;
  std::optional<int32_t> maxPriorityThreshold;  // BorrowWithinCohort.MaxPriorityThreshold
  uint8_t whenCanBorrow = KB_FUNG_MAY_STOP_SEARCH, whenCanPreempt = KB_FUNG_TRY_NEXT_FLAVOR, preference = KB_PREF_UNSET;
  uint8_t queueingStrategy = KB_QUEUE_BEST_EFFORT_FIFO;
  double fairWeight = 1.0;
  int64_t allocatableResourceGeneration = 1;
};
struct Cohort {
  std::string name, parent;
  std::vector<ResourceGroup> resourceGroups;
  double fairWeight = 1.0;
};

struct PodSet {  // PodSetResources: Requests are per pod here, multiplied by Count when flattened (workload.go:567-598)
  std::string name = "main";
  int32_t count = 1;
  std::optional<int32_t> minCount;
  std::map<std::string, int64_t> requests;
  std::optional<std::vector<std::string>> eligibleFlavors;  // checkFlavorForPodSets result (taints / affinity); nullopt = all
  std::map<std::string, int8_t> lastTriedFlavorIdx;         // LastState.LastTriedFlavorIdx per resource
};
struct FlavorUsage { std::string flavor, resource; int64_t quantity; };
struct WorkloadInfo {
  std::string key;           // namespace/name
  std::string clusterQueue;
  int32_t priority = 0;
  int64_t queueOrderTimestampNs = 0;  // Ordering.GetQueueOrderTimestamp (workload.go:1174-1193)
  int64_t uid = 0;                    // rank consistent with the UID string order
  std::vector<PodSet> podSets;
  int64_t lastAssignmentGeneration = -1;  // LastAssignment.ClusterQueueGeneration or -1
  // admitted workloads only
  std::vector<FlavorUsage> usage;              // Info.FlavorResourceUsage()
  std::optional<int64_t> quotaReservedNs;      // QuotaReserved.LastTransitionTime
  bool evicted = false;
};

struct Snapshot {  // what cache.Snapshot() hands to the cycle
  std::vector<ClusterQueue> clusterQueues;
  std::vector<Cohort> cohorts;              // cohorts only named as parents are created implicitly (hierarchy manager)
  std::vector<WorkloadInfo> admitted;       // ClusterQueueSnapshot.Workloads of every ClusterQueue
  std::vector<std::string> resourceFlavors; // optional explicit flavor order
  int64_t generation = 0;                   // bump on any ClusterQueue / Cohort spec change (kb_snapshot.static_generation)
};

struct PodSetAssignment {
  std::string name;
  int32_t count = 0;
  std::map<std::string, std::string> flavors;  // resource -> flavor
  std::map<std::string, int> modes;            // resource -> KB_MODE_*
  std::map<std::string, int> triedFlavorIdx;   // resource -> TriedFlavorIdx
};
struct Target { std::string key; int reason; };
struct Entry {
  std::string key;
  int decision = KB_DEC_NOFIT;  // KB_DEC_*
  int mode = KB_MODE_NOFIT;     // Assignment.RepresentativeMode()
  int borrowing = 0;            // Assignment.Borrowing
  int commitRank = -1;          // position in the iterator order inside the root cohort
  std::vector<PodSetAssignment> podSets;
  std::vector<Target> targets;
};

// queues.Heads (pkg/cache/queue/manager.go:770-794): the first workload of every ClusterQueue in queueOrderingFunc
// order (cluster_queue.go:636-685: higher priority, then queue-order timestamp, then UID).  One workload per
// ClusterQueue, in first-appearance order of the ClusterQueues.
inline std::vector<WorkloadInfo> selectHeads(const std::vector<WorkloadInfo> &pending) {
  auto before = [](const WorkloadInfo &a, const WorkloadInfo &b) {
    if (a.priority != b.priority) return a.priority > b.priority;
    if (a.queueOrderTimestampNs != b.queueOrderTimestampNs) return a.queueOrderTimestampNs < b.queueOrderTimestampNs;
    return a.uid < b.uid;
  };
  std::vector<WorkloadInfo> heads;
  for (const WorkloadInfo &w : pending) {
    auto it = std::find_if(heads.begin(), heads.end(), [&](const WorkloadInfo &h) { return h.clusterQueue == w.clusterQueue; });
    if (it == heads.end()) heads.push_back(w);
    else if (before(w, *it)) *it = w;
  }
  return heads;
}

// Owns the SoA buffers a kb_snapshot points into.
class FlatSnapshot {
 public:
  kb_snapshot s{};
  std::vector<std::string> cqNames, cohortNames, flavors, resources, pendingKeys, admittedKeys;

  int fr(const std::string &flavor, const std::string &resource) const {
    return index(flavors, flavor) * (int)resources.size() + index(resources, resource);
  }
  static int index(const std::vector<std::string> &v, const std::string &x) {
    auto it = std::find(v.begin(), v.end(), x);
    if (it == v.end()) throw Error(KB_ERR_INVALID, "unknown name: " + x);
    return (int)(it - v.begin());
  }

  FlatSnapshot(const Snapshot &snap, const std::vector<WorkloadInfo> &heads, uint32_t flags, int64_t nowNs) {
    // implicit cohorts (hierarchy.Manager, manager.go:80-100)
    std::vector<Cohort> cohorts = snap.cohorts;
    auto known = [&](const std::string &n) { return std::any_of(cohorts.begin(), cohorts.end(), [&](const Cohort &c) { return c.name == n; }); };
    for (auto &c : snap.clusterQueues) if (!c.cohort.empty() && !known(c.cohort)) { Cohort n; n.name = c.cohort; cohorts.push_back(n); }
    for (size_t i = 0; i < cohorts.size(); i++) if (!cohorts[i].parent.empty() && !known(cohorts[i].parent)) { Cohort n; n.name = cohorts[i].parent; cohorts.push_back(n); }
    flavors = snap.resourceFlavors;
    std::vector<std::string> res;
    auto addf = [&](const std::string &f) { if (std::find(flavors.begin(), flavors.end(), f) == flavors.end()) flavors.push_back(f); };
    auto addr = [&](const std::string &r) { if (std::find(res.begin(), res.end(), r) == res.end()) res.push_back(r); };
    auto scan = [&](const std::vector<ResourceGroup> &rgs) { for (auto &rg : rgs) for (auto &fq : rg) { addf(fq.flavor); for (auto &q : fq.resources) addr(q.resource); } };
    for (auto &c : snap.clusterQueues) scan(c.resourceGroups);
    for (auto &c : cohorts) scan(c.resourceGroups);
    for (auto &w : heads) for (auto &p : w.podSets) for (auto &kv : p.requests) addr(kv.first);
    for (auto &w : snap.admitted) for (auto &u : w.usage) { addr(u.resource); addf(u.flavor); }
    std::sort(res.begin(), res.end());  // resource index order == name order (DRS tie-break fair_sharing.go:149)
    if (res.empty()) res.push_back("cpu");
    if (flavors.empty()) flavors.push_back("default");
    resources = res;
    for (auto &c : snap.clusterQueues) cqNames.push_back(c.name);
    for (auto &c : cohorts) cohortNames.push_back(c.name);
    const int Q = (int)cqNames.size(), C = (int)cohortNames.size(), N = Q + C, F = (int)flavors.size(), R = (int)resources.size(), FR = F * R;
    if (R > KB_MAX_RESOURCES) throw Error(KB_ERR_INVALID, "too many resources");

    parent_.assign(N, -1); weight_.assign(N, 1.0);
    nominal_.assign((size_t)N * FR, 0); blimit_.assign((size_t)N * FR, KB_NO_LIMIT); llimit_.assign((size_t)N * FR, KB_NO_LIMIT);
    rgStart_.assign(1, 0); rgFlStart_.assign(1, 0);
    auto quotas = [&](int n, const std::vector<ResourceGroup> &rgs, bool isCq) {
      for (auto &rg : rgs) {
        uint32_t mask = 0;
        for (auto &fq : rg)
          for (auto &q : fq.resources) {
            size_t c = (size_t)n * FR + fr(fq.flavor, q.resource);
            nominal_[c] = q.nominal;
            if (q.borrowingLimit) blimit_[c] = *q.borrowingLimit;
            if (q.lendingLimit) llimit_[c] = *q.lendingLimit;
            mask |= 1u << index(resources, q.resource);
          }
        if (isCq) {
          rgMask_.push_back(mask);
          for (auto &fq : rg) rgFl_.push_back(index(flavors, fq.flavor));
          rgFlStart_.push_back((int32_t)rgFl_.size());
        }
      }
      if (isCq) rgStart_.push_back((int32_t)rgMask_.size());
    };
    for (int i = 0; i < Q; i++) {
      const ClusterQueue &c = snap.clusterQueues[i];
      if (!c.cohort.empty()) parent_[i] = Q + index(cohortNames, c.cohort);
      weight_[i] = c.fairWeight;
      quotas(i, c.resourceGroups, true);
      within_.push_back(c.withinClusterQueue); reclaim_.push_back(c.reclaimWithinCohort); bwc_.push_back(c.borrowWithinCohort);
      hasThr_.push_back(c.maxPriorityThreshold ? 1 : 0); thr_.push_back(c.maxPriorityThreshold.value_or(0));
      wcb_.push_back(c.whenCanBorrow); wcp_.push_back(c.whenCanPreempt); pref_.push_back(c.preference);
      strat_.push_back(c.queueingStrategy); gen_.push_back(c.allocatableResourceGeneration);
    }
    for (int i = 0; i < C; i++) {
      if (!cohorts[i].parent.empty()) parent_[Q + i] = Q + index(cohortNames, cohorts[i].parent);
      weight_[Q + i] = cohorts[i].fairWeight;
      quotas(Q + i, cohorts[i].resourceGroups, false);
    }
    // admitted workloads; ClusterQueue usage = sum of their usage (clusterqueue.go:535-563)
    cqUsage_.assign((size_t)std::max(1, Q) * FR, 0);
    admUseStart_.assign(1, 0);
    for (auto &w : snap.admitted) {
      int cq = index(cqNames, w.clusterQueue);
      admCq_.push_back(cq); admPrio_.push_back(w.priority); admTs_.push_back(w.queueOrderTimestampNs);
      admQr_.push_back(w.quotaReservedNs.value_or(INT64_MIN)); admUid_.push_back(w.uid); admEv_.push_back(w.evicted ? 1 : 0);
      std::map<int, int64_t> acc;
      for (auto &u : w.usage) acc[fr(u.flavor, u.resource)] += u.quantity;
      for (auto &kv : acc) { admFr_.push_back(kv.first); admQty_.push_back(kv.second); cqUsage_[(size_t)cq * FR + kv.first] += kv.second; }
      admUseStart_.push_back((int32_t)admFr_.size());
      admittedKeys.push_back(w.key);
    }
    // entries
    wlPsStart_.assign(1, 0);
    for (auto &w : heads) {
      wlCq_.push_back(index(cqNames, w.clusterQueue)); wlPrio_.push_back(w.priority); wlTs_.push_back(w.queueOrderTimestampNs);
      wlUid_.push_back(w.uid); wlGen_.push_back(w.lastAssignmentGeneration);
      for (auto &p : w.podSets) {
        size_t base = psReq_.size();
        psReq_.resize(base + R, 0); psLast_.resize(psLast_.size() + R, -1);
        uint32_t mask = 0;
        for (auto &kv : p.requests) { int r = index(resources, kv.first); psReq_[base + r] = kv.second * p.count; mask |= 1u << r; }
        for (auto &kv : p.lastTriedFlavorIdx) psLast_[psLast_.size() - R + index(resources, kv.first)] = kv.second;
        uint64_t ok = ~0ull;
        if (p.eligibleFlavors) { ok = 0; for (auto &f : *p.eligibleFlavors) ok |= 1ull << index(flavors, f); }
        psMask_.push_back(mask); psCount_.push_back(p.count); psMin_.push_back(p.minCount.value_or(-1)); psOk_.push_back(ok);
      }
      wlPsStart_.push_back((int32_t)psCount_.size());
      heads_.push_back((int32_t)pendingKeys.size());
      pendingKeys.push_back(w.key);
    }
    auto ptr = [](auto &v) { v.reserve(1); return v.data(); };
    s.n_cq = Q; s.n_cohort = C; s.n_flavor = F; s.n_resource = R; s.n_rg = (int32_t)rgMask_.size();
    s.n_wl = (int32_t)wlCq_.size(); s.n_podset = (int32_t)psCount_.size(); s.n_adm = (int32_t)admCq_.size();
    s.n_adm_use = (int32_t)admFr_.size(); s.n_heads = (int32_t)heads_.size();
    auto pods = std::find(resources.begin(), resources.end(), "pods");
    s.pods_resource = pods == resources.end() ? -1 : (int32_t)(pods - resources.begin());
    s.flags = flags; s.now_ns = nowNs; s.static_generation = snap.generation;
    s.parent = ptr(parent_); s.fair_weight = ptr(weight_); s.nominal = ptr(nominal_); s.borrow_limit = ptr(blimit_); s.lend_limit = ptr(llimit_);
    s.cq_usage = ptr(cqUsage_);
    s.cq_within_cq = ptr(within_); s.cq_reclaim_within = ptr(reclaim_); s.cq_borrow_within = ptr(bwc_);
    s.cq_has_bwc_threshold = ptr(hasThr_); s.cq_bwc_threshold = ptr(thr_);
    s.cq_when_can_borrow = ptr(wcb_); s.cq_when_can_preempt = ptr(wcp_); s.cq_preference = ptr(pref_); s.cq_strategy = ptr(strat_);
    s.cq_generation = ptr(gen_);
    s.cq_rg_start = ptr(rgStart_); s.rg_res_mask = ptr(rgMask_); s.rg_flavor_start = ptr(rgFlStart_); s.rg_flavors = ptr(rgFl_);
    s.wl_cq = ptr(wlCq_); s.wl_priority = ptr(wlPrio_); s.wl_ts = ptr(wlTs_); s.wl_uid = ptr(wlUid_); s.wl_last_gen = ptr(wlGen_);
    s.wl_ps_start = ptr(wlPsStart_);
    s.ps_req = ptr(psReq_); s.ps_req_mask = ptr(psMask_); s.ps_count = ptr(psCount_); s.ps_min_count = ptr(psMin_);
    s.ps_flavor_ok = ptr(psOk_); s.ps_last_tried = ptr(psLast_);
    s.adm_cq = ptr(admCq_); s.adm_priority = ptr(admPrio_); s.adm_ts = ptr(admTs_); s.adm_qr_ts = ptr(admQr_); s.adm_uid = ptr(admUid_);
    s.adm_evicted = ptr(admEv_); s.adm_use_start = ptr(admUseStart_); s.adm_use_fr = ptr(admFr_); s.adm_use_qty = ptr(admQty_);
    s.heads = ptr(heads_);
  }
  FlatSnapshot(const FlatSnapshot &) = delete;
  FlatSnapshot &operator=(const FlatSnapshot &) = delete;

 private:
  std::vector<int32_t> parent_, thr_, rgStart_, rgFlStart_, rgFl_, wlCq_, wlPrio_, wlPsStart_, psCount_, psMin_, admCq_, admPrio_, admUseStart_, admFr_, heads_;
  std::vector<double> weight_;
  std::vector<int64_t> nominal_, blimit_, llimit_, cqUsage_, gen_, wlTs_, wlUid_, wlGen_, psReq_, admTs_, admQr_, admUid_, admQty_;
  std::vector<uint8_t> within_, reclaim_, bwc_, hasThr_, wcb_, wcp_, pref_, strat_, admEv_;
  std::vector<uint32_t> rgMask_, psMask_;
  std::vector<uint64_t> psOk_;
  std::vector<int8_t> psLast_;
};

// One kb_handle (one device, one stream).  Not re-entrant, like the reference's single scheduling goroutine.
class Scheduler {
 public:
  explicit Scheduler(int device = 0, uint32_t flags = KB_FLAGS_DEFAULT) : flags_(flags) {
    kb_config cfg{};
    cfg.device = device;
    int32_t rc = kb_create(&cfg, &h_);
    if (rc != KB_OK) throw Error(rc, "kb_create failed (no CUDA device or library): the caller keeps the stock cycle");
  }
  ~Scheduler() { if (h_) kb_destroy(h_); }
  Scheduler(const Scheduler &) = delete;
  Scheduler &operator=(const Scheduler &) = delete;

  // Decision part of Scheduler.schedule (scheduler.go:245-405) for the given heads.
  std::vector<Entry> schedule(const Snapshot &snap, const std::vector<WorkloadInfo> &heads, int64_t nowNs) {
    FlatSnapshot f(snap, heads, flags_, nowNs);
    const kb_snapshot &s = f.s;
    const int H = s.n_heads, P = s.n_podset, R = s.n_resource;
    std::vector<uint8_t> decision(H + 1), mode(H + 1), reason;
    std::vector<int32_t> borrow(H + 1), rank(H + 1), psCount(P + 1), tgtStart(H + 2), tgtAdm;
    std::vector<int8_t> psFlavor((size_t)P * R + 1), psMode((size_t)P * R + 1), psTried((size_t)P * R + 1);
    int32_t cap = 4 * s.n_adm + 1024;
    kb_cycle_out out{};
    for (;;) {
      tgtAdm.assign(cap, 0); reason.assign(cap, 0);
      out.decision = decision.data(); out.mode = mode.data(); out.borrow = borrow.data(); out.commit_rank = rank.data();
      out.ps_flavor = psFlavor.data(); out.ps_res_mode = psMode.data(); out.ps_tried_idx = psTried.data(); out.ps_count = psCount.data();
      out.tgt_start = tgtStart.data(); out.tgt_adm = tgtAdm.data(); out.tgt_reason = reason.data(); out.tgt_capacity = cap;
      out.node_usage = nullptr;
      int32_t rc = kb_run_cycle(h_, &s, &out);
      if (rc == KB_ERR_CAPACITY) { cap *= 4; continue; }
      if (rc != KB_OK) throw Error(rc, kb_last_error(h_));
      break;
    }
    std::vector<Entry> entries(H);
    for (int e = 0; e < H; e++) {
      const WorkloadInfo &w = heads[e];
      Entry &en = entries[e];
      en.key = w.key; en.decision = decision[e]; en.mode = mode[e]; en.borrowing = borrow[e]; en.commitRank = rank[e];
      int row = s.wl_ps_start[e];  // heads are flattened in order, one pending record per head
      for (const PodSet &p : w.podSets) {
        PodSetAssignment a;
        a.name = p.name; a.count = psCount[row];
        for (int r = 0; r < R; r++) {
          int fl = psFlavor[(size_t)row * R + r];
          if (fl < 0) continue;
          a.flavors[f.resources[r]] = f.flavors[fl];
          a.modes[f.resources[r]] = psMode[(size_t)row * R + r];
          a.triedFlavorIdx[f.resources[r]] = psTried[(size_t)row * R + r];
        }
        en.podSets.push_back(std::move(a));
        row++;
      }
      for (int k = tgtStart[e]; k < tgtStart[e + 1]; k++) en.targets.push_back({f.admittedKeys[tgtAdm[k]], reason[k]});
    }
    return entries;
  }

 private:
  kb_handle *h_ = nullptr;
  uint32_t flags_;
};

}  // namespace kb
