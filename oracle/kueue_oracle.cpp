// kueue_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Single-threaded CPU restatement of Kueue's scheduling cycle, used only as
// the parity checker (tests/, __graft_entry__.smoke()) and as bench.py's
// cpu_baseline / --impl reference arm.  Nothing under kueue_b200/ may link,
// import or call this file.
//
// Every function cites the reference file:line it restates (paths relative to
// /root/reference).  The restatement keeps the reference's control flow
// (recursive available(), per-head nominate, sequential admit loop, greedy
// remove/fill-back preemption) on dense arrays instead of Go maps.
//
// Pinning (parity pinned): checked against the reference's own table tests, transcribed under tests/golden/
// (scripts in tools/, the Go literals are parsed, never executed): TestAvailable, TestDominantResourceShare,
// TestAssignFlavors, TestReclaimBeforePriorityPreemption, TestHierarchical, TestSearch (podset reducer),
// TestPreemption, TestHierarchicalPreemptions, TestFairPreemptions, TestCandidatesOrdering,
// TestSatisfiesPreemptionPolicy, TestSchedule (whole cycle), TestEntryOrdering, TestEntryComparerLess,
// TestResourcesToReserve — tests/test_oracle_golden_*.py; DESIGN.md (c) has the case counts.
//
// Canonical tie-breaks where the reference is nondeterministic:
//   * classical iterator ties (scheduler.go:779 unstable sort, comparator 0 at
//     :816): break by the entry's position in the cycle's entry list (heads[]).
//   * child iteration (hierarchy/cohort.go:44-50 UnsortedList): child cohorts
//     ascending node index, then child CQs ascending index.
//   * resource iteration inside findFlavorForPodSets / assignFlavors (Go map
//     order, flavorassigner.go:639,814): ascending resource index; only
//     message strings depend on it in the reference.
//   * fairSharingIterator.getCq (fair_sharing_iterator.go:103-108): roots in
//     ascending index; cross-root order does not affect decisions.

#include "../include/kueue_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <utility>
#include <array>
#include <vector>

namespace {

typedef int64_t i64;

enum PMode { P_NOFIT = 0, P_NOCAND = 1, P_PREEMPT = 2, P_RECLAIM = 3, P_FIT = 4 };  // flavorassigner.go:399-407
enum Variant { V_NEVER = 0, V_WITHIN_CQ, V_HIER_RECLAIM, V_RECLAIM_NO_BORROW, V_RECLAIM_WHILE_BORROW };  // hierarchical_preemption.go:33-47

struct GranularMode {  // flavorassigner.go:384-395
  int pmode;
  int borrow;
};
static const int kMaxInt = std::numeric_limits<int>::max();

struct DRS {  // fair_sharing.go:43-50
  double fairWeight = 1.0;
  double unweightedRatio = 0.0;
  int dominantResource = -1;
  bool borrowing = false;
  bool isZero() const { return unweightedRatio == 0; }
  bool isWeightZero() const { return fairWeight == 0; }
  bool zeroWeightBorrows() const { return isWeightZero() && !isZero(); }  // :121-123
  double precise() const {                                              // :75-83
    if (isZero()) return 0.0;
    if (isWeightZero()) return INFINITY;
    return unweightedRatio / fairWeight;
  }
};
static DRS negativeDRS() {  // :53-55
  DRS d; d.unweightedRatio = -1; d.fairWeight = 1.0; return d;
}
static int cmpD(double a, double b) { return a < b ? -1 : (a > b ? 1 : 0); }
static int compareDRS(const DRS &a, const DRS &b) {  // :89-100
  if (a.zeroWeightBorrows() && b.zeroWeightBorrows()) return cmpD(a.unweightedRatio, b.unweightedRatio);
  if (a.zeroWeightBorrows()) return 1;
  if (b.zeroWeightBorrows()) return -1;
  return cmpD(a.precise(), b.precise());
}

struct UsageVec {  // resources.FlavorResourceQuantities as a small (fr, qty) list
  std::vector<std::pair<int, i64>> v;
  void add(int fr, i64 q) {
    for (auto &p : v) if (p.first == fr) { p.second += q; return; }
    v.emplace_back(fr, q);
  }
  i64 get(int fr) const {
    for (auto &p : v) if (p.first == fr) return p.second;
    return 0;
  }
};

struct PodSetAssign {  // flavorassigner.go:262-273
  int count = 0;
  int8_t flavor[KB_MAX_RESOURCES];
  int8_t mode[KB_MAX_RESOURCES];
  int8_t tried[KB_MAX_RESOURCES];
  int borrow[KB_MAX_RESOURCES];
  bool hasReasons = false;  // !Status.IsFit()
  int nFlavors = 0;
  PodSetAssign() {
    for (int r = 0; r < KB_MAX_RESOURCES; r++) { flavor[r] = -1; mode[r] = -1; tried[r] = -1; borrow[r] = 0; }
  }
  int repMode() const {  // :277-295
    if (!hasReasons) return KB_MODE_FIT;
    if (nFlavors == 0) return KB_MODE_NOFIT;
    int m = KB_MODE_FIT;
    for (int r = 0; r < KB_MAX_RESOURCES; r++) if (flavor[r] >= 0 && mode[r] < m) m = mode[r];
    return m;
  }
};

struct Assignment {  // flavorassigner.go:44-73
  std::vector<PodSetAssign> ps;
  UsageVec usage;  // Usage.Quota
  int borrowing = 0;
  int repMode() const {  // :147-164
    if (ps.empty()) return KB_MODE_NOFIT;
    int m = KB_MODE_FIT;
    for (auto &p : ps) m = std::min(m, p.repMode());
    return m;
  }
};

struct Target { int adm; int reason; };

struct Entry {
  int wl;
  Assignment a;
  std::vector<Target> targets;
  int decision = KB_DEC_NOFIT;
  int rank = -1;
};

struct Preemptor {  // what preemption needs to know about the incoming workload
  int cq; int priority; i64 ts;
};

class Oracle {
 public:
  const kb_snapshot &s;
  int Q, C, N, F, R, FR;
  std::vector<i64> subtree, usage;  // [N][FR]
  std::vector<std::vector<int>> childCohorts, childCqs;
  std::vector<int> height;
  std::vector<std::vector<int>> cqAdm;  // ClusterQueueSnapshot.Workloads
  bool fair;
  const int8_t *stubMode = nullptr;   // testOracle (flavorassigner_test.go:145-158): per-FR stub result
  const int32_t *stubBorrow = nullptr;
  bool useStub = false;

  explicit Oracle(const kb_snapshot &snap) : s(snap) {
    Q = s.n_cq; C = s.n_cohort; N = Q + C; F = s.n_flavor; R = s.n_resource; FR = F * R;
    fair = (s.flags & KB_F_FAIR_SHARING) != 0;
    subtree.assign((size_t)N * FR, 0);
    usage.assign((size_t)N * FR, 0);
    childCohorts.resize(N); childCqs.resize(N);
    for (int n = 0; n < N; n++) {
      int p = s.parent[n];
      if (p >= 0) { if (n < Q) childCqs[p].push_back(n); else childCohorts[p].push_back(n); }
    }
    for (int q = 0; q < Q; q++)
      for (int fr = 0; fr < FR; fr++) {
        subtree[(size_t)q * FR + fr] = s.nominal[(size_t)q * FR + fr];  // updateClusterQueueResourceNode resource_node.go:160-166
        usage[(size_t)q * FR + fr] = s.cq_usage[(size_t)q * FR + fr];
      }
    height.assign(N, 0);
    for (int n = Q; n < N; n++) if (s.parent[n] < 0) updateCohortResourceNode(n);
    for (int n = Q; n < N; n++) height[n] = getNodeHeight(n);
    cqAdm.resize(Q);
    for (int a = 0; a < s.n_adm; a++) cqAdm[s.adm_cq[a]].push_back(a);
  }

  // ---- resource_node.go -------------------------------------------------
  bool hasParent(int n) const { return s.parent[n] >= 0; }
  i64 &U(int n, int fr) { return usage[(size_t)n * FR + fr]; }
  i64 Sub(int n, int fr) const { return subtree[(size_t)n * FR + fr]; }
  i64 Nominal(int n, int fr) const { return s.nominal[(size_t)n * FR + fr]; }
  i64 localQuota(int n, int fr) const {  // :66-71
    i64 ll = s.lend_limit[(size_t)n * FR + fr];
    if (ll != KB_NO_LIMIT) return std::max<i64>(0, Sub(n, fr) - ll);
    return 0;
  }
  i64 localAvailable(int n, int fr) {  // :91-93
    return std::max<i64>(0, localQuota(n, fr) - U(n, fr));
  }
  i64 available(int n, int fr) {  // :104-118
    if (!hasParent(n)) return Sub(n, fr) - U(n, fr);
    i64 parentAvailable = available(s.parent[n], fr);
    i64 bl = s.borrow_limit[(size_t)n * FR + fr];
    if (bl != KB_NO_LIMIT) {
      i64 storedInParent = Sub(n, fr) - localQuota(n, fr);
      i64 usedInParent = std::max<i64>(0, U(n, fr) - localQuota(n, fr));
      i64 withMaxFromParent = storedInParent - usedInParent + bl;
      parentAvailable = std::min(withMaxFromParent, parentAvailable);
    }
    return localAvailable(n, fr) + parentAvailable;
  }
  i64 potentialAvailable(int n, int fr) {  // :122-133
    if (!hasParent(n)) return Sub(n, fr);
    i64 av = localQuota(n, fr) + potentialAvailable(s.parent[n], fr);
    i64 bl = s.borrow_limit[(size_t)n * FR + fr];
    if (bl != KB_NO_LIMIT) av = std::min(Sub(n, fr) + bl, av);
    return av;
  }
  void addUsage(int n, int fr, i64 val) {  // :137-145
    i64 la = localAvailable(n, fr);
    U(n, fr) += val;
    if (hasParent(n) && val > la) addUsage(s.parent[n], fr, val - la);
  }
  void removeUsage(int n, int fr, i64 val) {  // :149-158
    i64 usageStoredInParent = U(n, fr) - localQuota(n, fr);
    U(n, fr) -= val;
    if (usageStoredInParent <= 0 || !hasParent(n)) return;
    removeUsage(s.parent[n], fr, std::min(val, usageStoredInParent));
  }
  void updateCohortResourceNode(int c) {  // :183-198
    for (int fr = 0; fr < FR; fr++) { subtree[(size_t)c * FR + fr] = Nominal(c, fr); usage[(size_t)c * FR + fr] = 0; }
    for (int ch : childCohorts[c]) { updateCohortResourceNode(ch); accumulateFromChild(c, ch); }
    for (int ch : childCqs[c]) accumulateFromChild(c, ch);
  }
  void accumulateFromChild(int p, int ch) {  // :210-217
    for (int fr = 0; fr < FR; fr++) {
      subtree[(size_t)p * FR + fr] += Sub(ch, fr) - localQuota(ch, fr);
      usage[(size_t)p * FR + fr] += std::max<i64>(0, U(ch, fr) - localQuota(ch, fr));
    }
  }
  // clusterqueue_snapshot.go / cohort_snapshot.go
  i64 Available(int cq, int fr) { return std::max<i64>(0, available(cq, fr)); }  // :154-156
  bool borrowingWith(int n, int fr, i64 val) {  // cq :149-151 (Nominal), cohort_snapshot.go:90-92 (SubtreeQuota)
    if (n < Q) return U(n, fr) + val > Nominal(n, fr);
    return U(n, fr) + val > Sub(n, fr);
  }
  void addUsageVec(int cq, const UsageVec &u) { for (auto &p : u.v) addUsage(cq, p.first, p.second); }       // :94-99
  void removeUsageVec(int cq, const UsageVec &u) { for (auto &p : u.v) removeUsage(cq, p.first, p.second); } // :101-106
  bool fitsVec(int cq, const UsageVec &u) {  // :121-136
    for (auto &p : u.v) if (Available(cq, p.first) < p.second) return false;
    return true;
  }
  bool isWithinNominalInResources(int n, const std::vector<int> &frs) {  // resource_node.go:248-255
    for (int fr : frs) if (U(n, fr) > Sub(n, fr)) return false;
    return true;
  }
  // admitted workload usage: Snapshot.RemoveWorkload / AddWorkload snapshot.go:49-64
  void removeAdm(int a) {
    for (int k = s.adm_use_start[a]; k < s.adm_use_start[a + 1]; k++) removeUsage(s.adm_cq[a], s.adm_use_fr[k], s.adm_use_qty[k]);
  }
  void addAdm(int a) {
    for (int k = s.adm_use_start[a]; k < s.adm_use_start[a + 1]; k++) addUsage(s.adm_cq[a], s.adm_use_fr[k], s.adm_use_qty[k]);
  }
  int rootOf(int n) const { while (s.parent[n] >= 0) n = s.parent[n]; return n; }

  // ---- classical/hierarchical_preemption.go -----------------------------
  int getNodeHeight(int c) {  // :202-208
    int mh = std::min<int>((int)(childCohorts[c].size() + childCqs[c].size()), 1);
    for (int ch : childCohorts[c]) mh = std::max(mh, getNodeHeight(ch) + 1);
    return mh;
  }
  // FindHeightOfLowestSubtreeThatFits :214-227
  std::pair<int, bool> findHeight(int cq, int fr, i64 val) {
    if (!borrowingWith(cq, fr, val) || !hasParent(cq)) return {0, hasParent(cq)};
    i64 remaining = val - localAvailable(cq, fr);
    for (int t = s.parent[cq]; t >= 0; t = s.parent[t]) {
      if (!borrowingWith(t, fr, remaining)) return {height[t], hasParent(t)};
      remaining -= localAvailable(t, fr);
    }
    return {height[rootOf(cq)], false};
  }

  // Is fr a key of node n's SubtreeQuota map?  CQs: defined by their ResourceGroups
  // (resource.go:52-75); cohorts: own quota (approximated as any non-default cell) or
  // any child's (accumulateFromChild resource_node.go:210-213).  Only the wlReq test
  // path of dominantResourceShare can observe this.
  bool frDefined(int n, int fr) {
    if (n < Q) {
      int f = fr / R, r = fr % R;
      for (int g = s.cq_rg_start[n]; g < s.cq_rg_start[n + 1]; g++) {
        if (!(s.rg_res_mask[g] & (1u << r))) continue;
        for (int k = s.rg_flavor_start[g]; k < s.rg_flavor_start[g + 1]; k++) if (s.rg_flavors[k] == f) return true;
      }
      return false;
    }
    size_t i = (size_t)n * FR + fr;
    if (s.nominal[i] != 0 || s.borrow_limit[i] != KB_NO_LIMIT || s.lend_limit[i] != KB_NO_LIMIT) return true;
    for (int ch : childCohorts[n]) if (frDefined(ch, fr)) return true;
    for (int ch : childCqs[n]) if (frDefined(ch, fr)) return true;
    return false;
  }

  // ---- fair_sharing.go ----------------------------------------------------
  DRS dominantResourceShare(int n, const i64 *wlReq = nullptr) {  // :126-156
    DRS drs; drs.fairWeight = s.fair_weight[n];
    if (!hasParent(n)) return drs;
    i64 borrowing[KB_MAX_RESOURCES]; bool any = false;
    for (int r = 0; r < R; r++) borrowing[r] = 0;
    for (int fr = 0; fr < FR; fr++) {
      // wlReq only counts on FlavorResources that are keys of the node's SubtreeQuota
      // map (`for fr, quota := range SubtreeQuota`, :133); the scheduler always passes nil.
      i64 amountBorrowed = ((wlReq && frDefined(n, fr)) ? wlReq[fr] : 0) + U(n, fr) - Sub(n, fr);
      if (amountBorrowed > 0) { borrowing[fr % R] += amountBorrowed; any = true; }
    }
    if (!any) return drs;
    drs.borrowing = true;
    i64 lendable[KB_MAX_RESOURCES];  // calculateLendable :160-174
    for (int r = 0; r < R; r++) lendable[r] = 0;
    for (int fr = 0; fr < FR; fr++) lendable[fr % R] += potentialAvailable(s.parent[n], fr);
    for (int r = 0; r < R; r++) {
      i64 b = borrowing[r];
      if (b <= 0) continue;
      i64 lr = lendable[r];
      if (lr > 0) {
        double ratio = (double)b * 1000.0 / (double)lr;
        if (ratio > drs.unweightedRatio || (ratio == drs.unweightedRatio && r < drs.dominantResource)) {
          drs.unweightedRatio = ratio; drs.dominantResource = r;
        }
      }
    }
    return drs;
  }
  static i64 roundedWeightedShare(const DRS &d) {  // :110-118
    if (d.zeroWeightBorrows()) return std::numeric_limits<i64>::max();
    return (i64)std::ceil(d.precise());
  }

  // ---- preemption/common ---------------------------------------------------
  bool satisfiesPreemptionPolicy(const Preemptor &p, int a, int policy) {  // preemption_policy.go:30-48
    bool lower = p.priority > s.adm_priority[a];
    if (policy == KB_POLICY_LOWER_PRIORITY) return lower;
    if (policy == KB_POLICY_LOWER_OR_NEWER_EQUAL_PRIORITY) {
      bool newerEq = (p.priority == s.adm_priority[a]) && p.ts < s.adm_ts[a];
      if (newerEq && (s.flags & KB_F_TS_PREEMPTION_BUFFER)) newerEq = s.adm_ts[a] - p.ts > 300ll * 1000000000ll;  // timestampPreemptionBuffer :28
      return lower || newerEq;
    }
    return policy == KB_POLICY_ANY;
  }
  i64 quotaReservationTime(int a) const {  // ordering.go:93-100
    return s.adm_qr_ts[a] == INT64_MIN ? s.now_ns : s.adm_qr_ts[a];
  }
  // CandidatesOrdering ordering.go:41-100 (AdmissionFairSharing branch not modelled)
  int candidatesOrdering(int a, int b, int cq) {
    bool ea = s.adm_evicted[a], eb = s.adm_evicted[b];
    if (ea != eb) return ea ? -1 : 1;  // CompareBool: true first
    bool aIn = s.adm_cq[a] == cq, bIn = s.adm_cq[b] == cq;
    if (aIn != bIn) return bIn ? -1 : 1;  // CompareBool(b==cq, a==cq)
    if (s.adm_priority[a] != s.adm_priority[b]) return s.adm_priority[a] < s.adm_priority[b] ? -1 : 1;
    i64 ta = quotaReservationTime(a), tb = quotaReservationTime(b);
    if (ta != tb) return tb < ta ? -1 : 1;  // newer first
    if (s.adm_uid[a] != s.adm_uid[b]) return s.adm_uid[a] < s.adm_uid[b] ? -1 : 1;
    return 0;
  }
  bool workloadUsesResources(int a, const std::vector<int> &frs) {  // candidate_generator.go:52-61
    for (int k = s.adm_use_start[a]; k < s.adm_use_start[a + 1]; k++)
      for (int fr : frs) if (s.adm_use_fr[k] == fr) return true;
    return false;
  }

  // ---- preemption.go ------------------------------------------------------
  struct PCtx {
    Preemptor p;
    UsageVec workloadUsage;
    std::vector<int> frsNeed;
  };
  bool workloadFits(PCtx &c, bool allowBorrowing) {  // :550-561
    for (auto &pr : c.workloadUsage.v) {
      if (!allowBorrowing && borrowingWith(c.p.cq, pr.first, pr.second)) return false;
      if (pr.second > Available(c.p.cq, pr.first)) return false;
    }
    return true;
  }
  bool workloadFitsForFairSharing(PCtx &c) {  // :567-572
    removeUsageVec(c.p.cq, c.workloadUsage);
    bool r = workloadFits(c, true);
    addUsageVec(c.p.cq, c.workloadUsage);
    return r;
  }
  bool queueUnderNominal(PCtx &c) {  // :577-584
    for (int fr : c.frsNeed) if (U(c.p.cq, fr) >= Nominal(c.p.cq, fr)) return false;
    return true;
  }
  bool queueWithinNominal(PCtx &c) {  // :591-598
    for (int fr : c.frsNeed) if (borrowingWith(c.p.cq, fr, 0)) return false;
    return true;
  }
  void fillBackWorkloads(PCtx &c, std::vector<Target> &targets, bool allowBorrowing) {  // :295-308
    for (int i = (int)targets.size() - 2; i >= 0; i--) {
      addAdm(targets[i].adm);
      if (workloadFits(c, allowBorrowing)) {
        targets[i] = targets.back();
        targets.pop_back();
      } else {
        removeAdm(targets[i].adm);
      }
    }
  }
  void restoreSnapshot(const std::vector<Target> &t) { for (auto &x : t) addAdm(x.adm); }  // :310-314

  // classical -----------------------------------------------------------------
  struct Cand { int adm; int lca; int variant; };
  bool borrowWithinCohortForbidden(int cq) {  // hierarchical_preemption.go:72-78
    return s.cq_borrow_within[cq] == KB_POLICY_NEVER;
  }
  int classifyPreemptionVariant(PCtx &c, int a, bool hierAdv) {  // :82-114
    if (!workloadUsesResources(a, c.frsNeed)) return V_NEVER;
    bool same = s.adm_cq[a] == c.p.cq;
    int policy = same ? s.cq_within_cq[c.p.cq] : s.cq_reclaim_within[c.p.cq];
    if (!satisfiesPreemptionPolicy(c.p, a, policy)) return V_NEVER;
    if (same) return V_WITHIN_CQ;
    if (hierAdv) return V_HIER_RECLAIM;
    if (borrowWithinCohortForbidden(c.p.cq)) return V_RECLAIM_NO_BORROW;
    int candP = s.adm_priority[a], inP = c.p.priority;
    // isAboveBorrowingThreshold :116-124
    bool above;
    if (candP >= inP) above = true;
    else if (!s.cq_has_bwc_threshold[c.p.cq]) above = false;
    else above = candP > s.cq_bwc_threshold[c.p.cq];
    return above ? V_RECLAIM_NO_BORROW : V_RECLAIM_WHILE_BORROW;
  }
  void getCandidatesFromCQ(PCtx &c, int cq, int lca, bool hierAdv, std::vector<Cand> &out) {  // :133-149
    for (int a : cqAdm[cq]) {
      int v = classifyPreemptionVariant(c, a, hierAdv);
      if (v == V_NEVER) continue;
      out.push_back({a, lca, v});
    }
  }
  void collectCandidatesInSubtree(PCtx &c, int cur, int subtreeRoot, int skip, bool hierAdv, std::vector<Cand> &out) {  // :181-199
    for (int ch : childCohorts[cur]) {
      if (ch == skip) continue;
      if (isWithinNominalInResources(ch, c.frsNeed)) continue;
      collectCandidatesInSubtree(c, ch, subtreeRoot, skip, hierAdv, out);
    }
    for (int cq : childCqs[cur]) {
      if (cq == c.p.cq) continue;
      if (!isWithinNominalInResources(cq, c.frsNeed)) getCandidatesFromCQ(c, cq, subtreeRoot, hierAdv, out);
    }
  }
  // QuantitiesFitInQuota resource_node.go:234-244
  bool quantitiesFitInQuota(int n, UsageVec &req) {
    bool fits = true;
    for (auto &p : req.v) {
      if (U(n, p.first) + p.second > Sub(n, p.first)) fits = false;
      p.second = std::max<i64>(0, p.second - localAvailable(n, p.first));
    }
    return fits;
  }
  std::vector<Target> classicalPreemptions(PCtx &c) {  // preemption.go:238-293
    int cq = c.p.cq;
    // NewCandidateIterator candidate_generator.go:77-121
    std::vector<Cand> sameQ, hier, prio;
    if (s.cq_within_cq[cq] != KB_POLICY_NEVER) getCandidatesFromCQ(c, cq, -1, false, sameQ);  // collectSameQueueCandidates :126-131
    if (hasParent(cq) && s.cq_reclaim_within[cq] != KB_POLICY_NEVER) {  // collectCandidatesForHierarchicalReclaim :151-177
      UsageVec remaining = c.workloadUsage;
      bool hierAdv = quantitiesFitInQuota(cq, remaining);
      int prev = -1;
      for (int cur = s.parent[cq]; cur >= 0; cur = s.parent[cur]) {
        collectCandidatesInSubtree(c, cur, cur, prev, hierAdv, hierAdv ? hier : prio);
        bool fits = quantitiesFitInQuota(cur, remaining);
        hierAdv = hierAdv || fits;
        prev = cur;
      }
    }
    auto cmp = [&](const Cand &a, const Cand &b) { return candidatesOrdering(a.adm, b.adm, cq) < 0; };
    std::sort(sameQ.begin(), sameQ.end(), cmp);
    std::sort(prio.begin(), prio.end(), cmp);
    std::sort(hier.begin(), hier.end(), cmp);
    std::vector<Cand> all;
    auto appendEv = [&](std::vector<Cand> &v, bool ev) { for (auto &x : v) if ((bool)s.adm_evicted[x.adm] == ev) all.push_back(x); };
    appendEv(hier, true); appendEv(prio, true); appendEv(sameQ, true);
    appendEv(hier, false); appendEv(prio, false); appendEv(sameQ, false);
    bool noOther = hier.empty() && prio.empty();
    bool noHier = hier.empty();
    bool forbidden = borrowWithinCohortForbidden(cq);
    bool opts[2]; int nopts;
    if (noOther || (forbidden && !queueUnderNominal(c))) { opts[0] = true; nopts = 1; }       // :266-267
    else if (forbidden && noHier) { opts[0] = false; opts[1] = true; nopts = 2; }                // :268-269
    else { opts[0] = true; opts[1] = false; nopts = 2; }                                         // :270-271
    for (int oi = 0; oi < nopts; oi++) {
      bool borrow = opts[oi];
      std::vector<Target> targets;
      for (size_t i = 0; i < all.size(); i++) {
        if (!candidateIsValid(c, all[i], borrow)) continue;  // Next() candidate_generator.go:125-135
        removeAdm(all[i].adm);
        targets.push_back({all[i].adm, variantReason(all[i].variant)});
        if (workloadFits(c, borrow)) {
          fillBackWorkloads(c, targets, borrow);
          restoreSnapshot(targets);
          return targets;
        }
      }
      restoreSnapshot(targets);
    }
    return {};
  }
  static int variantReason(int v) {  // hierarchical_preemption.go:49-61
    switch (v) {
      case V_WITHIN_CQ: return KB_REASON_IN_CLUSTER_QUEUE;
      case V_HIER_RECLAIM: return KB_REASON_IN_COHORT_RECLAMATION;
      case V_RECLAIM_WHILE_BORROW: return KB_REASON_IN_COHORT_RECLAIM_WHILE_BORROWING;
      case V_RECLAIM_NO_BORROW: return KB_REASON_IN_COHORT_RECLAMATION;
    }
    return 0;
  }
  bool candidateIsValid(PCtx &c, const Cand &cand, bool borrow) {  // candidate_generator.go:140-162
    int ccq = s.adm_cq[cand.adm];
    if (c.p.cq == ccq) return true;
    if (borrow && cand.variant == V_RECLAIM_NO_BORROW) return false;
    if (isWithinNominalInResources(ccq, c.frsNeed)) return false;
    for (int n = s.parent[ccq]; n >= 0; n = s.parent[n]) {
      if (n == cand.lca) break;
      if (isWithinNominalInResources(n, c.frsNeed)) return false;
    }
    return true;
  }

  // fair ------------------------------------------------------------------------
  bool cqIsBorrowing(int cq, const std::vector<int> &frs) {  // preemption.go:535-545
    if (!hasParent(cq)) return false;
    for (int fr : frs) if (borrowingWith(cq, fr, 0)) return true;
    return false;
  }
  void subtreeClusterQueues(int c, std::vector<int> &out) {  // cohort_snapshot.go:55-66
    for (int q : childCqs[c]) out.push_back(q);
    for (int ch : childCohorts[c]) subtreeClusterQueues(ch, out);
  }
  std::vector<int> findCandidates(PCtx &c) {  // preemption.go:514-533
    std::vector<int> cand;
    int cq = c.p.cq;
    auto forPolicy = [&](int fromCq, int policy) {  // findCandidatesForPolicy :492-509
      for (int a : cqAdm[fromCq]) {
        if (!satisfiesPreemptionPolicy(c.p, a, policy)) continue;
        if (!workloadUsesResources(a, c.frsNeed)) continue;
        cand.push_back(a);
      }
    };
    if (s.cq_within_cq[cq] != KB_POLICY_NEVER) forPolicy(cq, s.cq_within_cq[cq]);
    if (hasParent(cq) && s.cq_reclaim_within[cq] != KB_POLICY_NEVER) {
      std::vector<int> cqs; subtreeClusterQueues(rootOf(cq), cqs);
      for (int o : cqs) {
        if (o == cq || !cqIsBorrowing(o, c.frsNeed)) continue;
        forPolicy(o, s.cq_reclaim_within[cq]);
      }
    }
    return cand;
  }
  // fairsharing/ordering.go
  struct FsOrdering {
    int preemptorCq;
    std::vector<char> isAncestor;             // preemptorAncestors
    std::vector<std::vector<int>> cqToTarget; // clusterQueueToTarget (front = next)
    std::vector<size_t> head;                 // pop index
    std::vector<char> prunedCq, prunedCohort;
  };
  FsOrdering makeOrdering(int cq, const std::vector<int> &cands) {  // :62-83
    FsOrdering t; t.preemptorCq = cq;
    t.isAncestor.assign(N, 0); t.prunedCq.assign(N, 0); t.prunedCohort.assign(N, 0);
    t.cqToTarget.resize(Q); t.head.assign(Q, 0);
    for (int n = s.parent[cq]; n >= 0; n = s.parent[n]) t.isAncestor[n] = 1;
    for (int a : cands) t.cqToTarget[s.adm_cq[a]].push_back(a);
    return t;
  }
  bool fsHasWorkload(FsOrdering &t, int cq) { return t.head[cq] < t.cqToTarget[cq].size(); }
  int fsPop(FsOrdering &t, int cq) { return t.cqToTarget[cq][t.head[cq]++]; }
  int nextTarget(FsOrdering &t, int cohort) {  // :141-208 ; returns cq or -1
    int highestCq = -1; DRS highestCqDrs = negativeDRS();
    for (int cq : childCqs[cohort]) {
      if (t.prunedCq[cq]) continue;
      DRS drs = dominantResourceShare(cq);
      if ((!drs.borrowing && cq != t.preemptorCq) || !fsHasWorkload(t, cq)) {
        t.prunedCq[cq] = 1;
      } else if (compareDRS(drs, highestCqDrs) == 0) {
        int newCand = t.cqToTarget[cq][t.head[cq]];
        int curCand = t.cqToTarget[highestCq][t.head[highestCq]];
        if (candidatesOrdering(newCand, curCand, t.preemptorCq) < 0) highestCq = cq;
      } else if (compareDRS(drs, highestCqDrs) == 1) {
        highestCqDrs = drs; highestCq = cq;
      }
    }
    int highestCohort = -1; DRS highestCohortDrs = negativeDRS();
    for (int ch : childCohorts[cohort]) {
      if (t.prunedCohort[ch]) continue;
      DRS drs = dominantResourceShare(ch);
      if (!drs.borrowing && !t.isAncestor[ch]) t.prunedCohort[ch] = 1;
      else if (compareDRS(drs, highestCohortDrs) >= 0) { highestCohortDrs = drs; highestCohort = ch; }
    }
    if (highestCohort < 0 && highestCq < 0) { t.prunedCohort[cohort] = 1; return -1; }
    if (compareDRS(highestCohortDrs, highestCqDrs) >= 0) return nextTarget(t, highestCohort);
    return highestCq;
  }
  // least_common_ancestor.go:27-58
  void almostLCAs(FsOrdering &t, int targetCq, int &preAl, int &tgtAl) {
    int lca = -1;
    for (int n = s.parent[targetCq]; n >= 0; n = s.parent[n]) if (t.isAncestor[n]) { lca = n; break; }
    auto al = [&](int cq) { int a = cq; for (int n = s.parent[cq]; n >= 0; n = s.parent[n]) { if (n == lca) return a; a = n; } return a; };
    preAl = al(t.preemptorCq); tgtAl = al(targetCq);
  }
  bool strategyS2a(const DRS &preNew, const DRS &tgtNew) { return compareDRS(preNew, tgtNew) <= 0; }  // strategy.go:41-43
  bool strategyS2b(const DRS &preNew, const DRS &tgtOld) { return compareDRS(preNew, tgtOld) < 0; }   // :46-48

  std::vector<Target> fairPreemptions(PCtx &c) {  // preemption.go:433-478
    int pcq = c.p.cq;
    std::vector<int> cands = findCandidates(c);
    if (cands.empty()) return {};
    std::sort(cands.begin(), cands.end(), [&](int a, int b) { return candidatesOrdering(a, b, pcq) < 0; });
    // parseStrategies :319-333
    bool s2a = s.flags & KB_F_FS_STRATEGY_S2A, s2b = s.flags & KB_F_FS_STRATEGY_S2B;
    int strat[2]; int nstrat = 0;  // 0 = S2-a, 1 = S2-b
    if (!s2a && !s2b) { strat[0] = 0; strat[1] = 1; nstrat = 2; }
    else if (s2a && s2b) { if (s.flags & KB_F_FS_STRATEGY_S2B_FIRST) { strat[0] = 1; strat[1] = 0; } else { strat[0] = 0; strat[1] = 1; } nstrat = 2; }
    else { strat[0] = s2a ? 0 : 1; nstrat = 1; }

    addUsageVec(pcq, c.workloadUsage);  // SimulateUsageAddition :446
    std::vector<Target> targets; std::vector<int> retry;
    bool fits = false;
    {  // runFirstFsStrategy :338-403
      FsOrdering t = makeOrdering(pcq, cands);
      bool withinNominal = (s.flags & KB_F_FS_PREEMPT_WITHIN_NOMINAL) && queueWithinNominal(c);
      auto step = [&](int tcq) -> bool {  // returns true when done (fits)
        if (tcq == pcq) {
          int a = fsPop(t, tcq); removeAdm(a); targets.push_back({a, KB_REASON_IN_CLUSTER_QUEUE});
          return workloadFitsForFairSharing(c);
        }
        if (withinNominal) {
          int a = fsPop(t, tcq); removeAdm(a); targets.push_back({a, KB_REASON_IN_COHORT_RECLAMATION});
          return workloadFitsForFairSharing(c);
        }
        int preAl, tgtAl; almostLCAs(t, tcq, preAl, tgtAl);  // ComputeShares target.go:53-56
        DRS preNew = dominantResourceShare(preAl), tgtOld = dominantResourceShare(tgtAl);
        while (fsHasWorkload(t, tcq)) {
          int a = fsPop(t, tcq);
          removeAdm(a);  // ComputeTargetShareAfterRemoval target.go:66-73
          int p2, t2; almostLCAs(t, tcq, p2, t2);
          DRS tgtNew = dominantResourceShare(t2);
          addAdm(a);
          bool ok = strat[0] == 0 ? strategyS2a(preNew, tgtNew) : strategyS2b(preNew, tgtOld);
          if (ok) {
            removeAdm(a); targets.push_back({a, KB_REASON_IN_COHORT_FAIR_SHARING});
            if (workloadFitsForFairSharing(c)) return true;
            break;
          } else retry.push_back(a);
        }
        return false;
      };
      if (!hasParent(pcq)) {  // ordering.go:92-103 (then the reference dereferences a nil parent; we stop)
        while (fsHasWorkload(t, pcq)) if (step(pcq)) { fits = true; break; }
      } else {
        int root = rootOf(pcq);
        while (!t.prunedCohort[root]) {
          int tcq = nextTarget(t, root);
          if (tcq < 0) continue;
          if (step(tcq)) { fits = true; break; }
        }
      }
    }
    if (!fits && nstrat > 1 && hasParent(pcq)) {  // runSecondFsStrategy :407-431
      FsOrdering t = makeOrdering(pcq, retry);
      int root = rootOf(pcq);
      while (!t.prunedCohort[root]) {
        int tcq = nextTarget(t, root);
        if (tcq < 0) continue;
        int preAl, tgtAl; almostLCAs(t, tcq, preAl, tgtAl);
        DRS preNew = dominantResourceShare(preAl), tgtOld = dominantResourceShare(tgtAl);
        if (strategyS2b(preNew, tgtOld)) {
          int a = fsPop(t, tcq); removeAdm(a); targets.push_back({a, KB_REASON_IN_COHORT_FAIR_SHARING});
          if (workloadFitsForFairSharing(c)) { fits = true; break; }
        }
        t.prunedCq[tcq] = 1;  // DropQueue
      }
    }
    removeUsageVec(pcq, c.workloadUsage);  // revertSimulation :459
    if (!fits) { restoreSnapshot(targets); return {}; }
    fillBackWorkloads(c, targets, true);
    restoreSnapshot(targets);
    return targets;
  }
  std::vector<Target> getTargets(PCtx &c) {  // preemption.go:148-153
    return fair ? fairPreemptions(c) : classicalPreemptions(c);
  }
  // preemption_oracle.go:41-71
  std::pair<int, int> simulatePreemption(const Preemptor &p, int fr, i64 quantity) {
    if (useStub) {  // testOracle.SimulatePreemption: table lookup, default (Preempt, 0)
      if (stubMode && stubMode[fr] >= 0) return {stubMode[fr], stubBorrow[fr]};
      return {P_PREEMPT, 0};
    }
    PCtx c; c.p = p; c.frsNeed = {fr}; c.workloadUsage.add(fr, quantity);
    std::vector<Target> cand = getTargets(c);
    if (cand.empty()) return {P_NOCAND, findHeight(p.cq, fr, quantity).first};  // :53-56
    for (auto &t : cand) removeAdm(t.adm);
    int borrowAfter = findHeight(p.cq, fr, quantity).first;
    for (auto &t : cand) addAdm(t.adm);
    for (auto &t : cand) if (s.adm_cq[t.adm] == p.cq) return {P_PREEMPT, borrowAfter};
    return {P_RECLAIM, borrowAfter};
  }

  // ---- flavorassigner.go ---------------------------------------------------
  bool isPreferred(const GranularMode &a, const GranularMode &b, int pref) {  // :410-441
    if (a.pmode == P_NOFIT) return false;
    if (b.pmode == P_NOFIT) return true;
    if (pref == KB_PREF_PREEMPTION_OVER_BORROWING) {
      if (a.borrow != b.borrow) return a.borrow < b.borrow;
      return a.pmode > b.pmode;
    }
    if (a.pmode != b.pmode) return a.pmode > b.pmode;
    return a.borrow < b.borrow;
  }
  bool shouldTryNextFlavor(const GranularMode &m, int cq) {  // :946-963
    if (m.pmode == P_NOFIT || m.pmode == P_NOCAND) return true;
    if ((m.pmode == P_PREEMPT || m.pmode == P_RECLAIM) && s.cq_when_can_preempt[cq] == KB_FUNG_TRY_NEXT_FLAVOR) return true;
    if (m.borrow != 0 && s.cq_when_can_borrow[cq] == KB_FUNG_TRY_NEXT_FLAVOR) return true;
    return false;
  }
  static int famode(int pm) { return pm == P_NOFIT ? KB_MODE_NOFIT : (pm == P_FIT ? KB_MODE_FIT : KB_MODE_PREEMPT); }  // :470-485
  bool canPreemptWhileBorrowing(int cq) {  // :1049-1052
    return s.cq_borrow_within[cq] != KB_POLICY_NEVER || (fair && s.cq_reclaim_within[cq] != KB_POLICY_NEVER);
  }
  // fitsResourceQuota :1017-1047
  std::pair<int, int> fitsResourceQuota(const Preemptor &p, int fr, i64 assumed, i64 request) {
    int cq = p.cq;
    i64 avail = Available(cq, fr);
    i64 maxCap = potentialAvailable(cq, fr);
    i64 val = assumed + request;
    if (val > maxCap) return {P_NOFIT, 0};
    auto hb = findHeight(cq, fr, val);
    if (val <= avail) return {P_FIT, hb.first};
    if (val <= Nominal(cq, fr) || hb.second || canPreemptWhileBorrowing(cq)) return simulatePreemption(p, fr, val);
    return {P_NOFIT, hb.first};
  }
  int rgByResource(int cq, int r) {  // clusterqueue_snapshot.go:67-74
    for (int g = s.cq_rg_start[cq]; g < s.cq_rg_start[cq + 1]; g++) if (s.rg_res_mask[g] & (1u << r)) return g;
    return -1;
  }
  // findFlavorForPodSets :762-897 for one podset or one PodSetGroup (psRow = psIDs[0], reqs / mask / okMask of the
  // whole group).  Returns false when no flavor could be assigned.  hasReasons <- status != nil.
  bool findFlavorForPodSet(int wl, int psRow, uint64_t okMask, const Preemptor &p, const i64 *reqs, uint32_t mask, int resName,
                           const UsageVec &assignmentUsage, bool useLast, PodSetAssign &out, bool &hasReasons) {
    int cq = p.cq;
    int g = rgByResource(cq, resName);
    if (g < 0) { hasReasons = true; return false; }  // :770-772
    uint32_t rgMask = s.rg_res_mask[g] & mask;        // filterRequestedResources :775
    int nfl = s.rg_flavor_start[g + 1] - s.rg_flavor_start[g];
    const int32_t *flv = s.rg_flavors + s.rg_flavor_start[g];
    int pref = s.cq_preference[cq];
    bool fung = s.flags & KB_F_FLAVOR_FUNGIBILITY;
    bool haveBest = false; GranularMode bestMode{P_NOFIT, kMaxInt};
    int8_t bFlavor[KB_MAX_RESOURCES], bMode[KB_MAX_RESOURCES]; int bBorrow[KB_MAX_RESOURCES];
    bool anyReason = false;
    int attempted = -1;
    int idx = 0;  // NextFlavorToTryForPodSetResource workload.go:178-191
    if (fung && useLast) idx = s.ps_last_tried[(size_t)psRow * R + resName] + 1;
    for (; idx < nfl; idx++) {
      attempted = idx;
      int f = flv[idx];
      if (!((okMask >> f) & 1)) { anyReason = true; continue; }  // checkFlavorForPodSets :798-806
      GranularMode rep{P_FIT, 0};
      int8_t aMode[KB_MAX_RESOURCES]; int aBorrow[KB_MAX_RESOURCES]; bool aSet[KB_MAX_RESOURCES];
      for (int r = 0; r < R; r++) aSet[r] = false;
      for (int r = 0; r < R; r++) {
        if (!(rgMask & (1u << r))) continue;
        int fr = f * R + r;
        auto pm = fitsResourceQuota(p, fr, assignmentUsage.get(fr), reqs[r]);  // :839
        if (pm.first != P_FIT) anyReason = true;
        GranularMode mode{pm.first, pm.second};
        if (isPreferred(rep, mode, pref)) rep = mode;  // :846-848
        if (rep.pmode == P_NOFIT) break;              // :849-852
        aSet[r] = true; aMode[r] = (int8_t)famode(pm.first); aBorrow[r] = pm.second;
      }
      auto take = [&]() {
        haveBest = true; bestMode = rep;
        for (int r = 0; r < R; r++) { bFlavor[r] = aSet[r] ? (int8_t)f : (int8_t)-1; bMode[r] = aSet[r] ? aMode[r] : (int8_t)-1; bBorrow[r] = aSet[r] ? aBorrow[r] : 0; }
      };
      if (fung) {  // :863-872
        if (!shouldTryNextFlavor(rep, cq)) { take(); break; }
        if (isPreferred(rep, bestMode, pref)) take();
      } else if (rep.pmode > bestMode.pmode) {  // :873-880
        take();
        if (bestMode.pmode == P_FIT) {
          for (int r = 0; r < R; r++) if (bFlavor[r] >= 0) { out.flavor[r] = bFlavor[r]; out.mode[r] = bMode[r]; out.borrow[r] = bBorrow[r]; out.tried[r] = 0; }
          return true;  // status nil
        }
      }
    }
    if (!haveBest) { hasReasons = hasReasons || anyReason; return false; }
    int tried = 0;
    if (fung) tried = (attempted == nfl - 1) ? -1 : attempted;  // :883-891
    for (int r = 0; r < R; r++) if (bFlavor[r] >= 0) { out.flavor[r] = bFlavor[r]; out.mode[r] = bMode[r]; out.borrow[r] = bBorrow[r]; out.tried[r] = (int8_t)tried; }
    if (fung && bestMode.pmode == P_FIT) return true;  // :892-894 status nil
    hasReasons = hasReasons || anyReason;
    return true;
  }
  // Assign :540-552 + assignFlavors :560-715 (TAS, workload slices not modelled)
  Assignment assign(int wl, const int32_t *counts) {
    int cq = s.wl_cq[wl];
    Preemptor p{cq, s.wl_priority[wl], s.wl_ts[wl]};
    bool useLast = s.wl_last_gen[wl] >= 0 && !(s.cq_generation[cq] > s.wl_last_gen[wl]);  // lastAssignmentOutdated :532-534
    Assignment a;
    int ps0 = s.wl_ps_start[wl], ps1 = s.wl_ps_start[wl + 1];
    bool coversPods = s.pods_resource >= 0 && rgByResource(cq, s.pods_resource) >= 0;
    int rowEnd;
    for (int row = ps0; row < ps1; row = rowEnd) {
      // one unit: a podset, or the adjacent podsets of one PodSetGroup (groupedRequests :613-631)
      rowEnd = row + 1;
      if (s.ps_group && s.ps_group[row] >= 0) while (rowEnd < ps1 && s.ps_group[rowEnd] == s.ps_group[row]) rowEnd++;
      const int nm = rowEnd - row;
      std::vector<std::array<i64, KB_MAX_RESOURCES>> mreq(nm);
      std::vector<uint32_t> mmask(nm); std::vector<int> mcount(nm);
      i64 reqs[KB_MAX_RESOURCES]; uint32_t mask = 0; uint64_t okMask = ~0ull;
      for (int r = 0; r < R; r++) reqs[r] = 0;
      for (int m = 0; m < nm; m++) {
        int mr = row + m;
        uint32_t mk = s.ps_req_mask[mr];
        int count = s.ps_count[mr];
        for (int r = 0; r < R; r++) mreq[m][r] = s.ps_req[(size_t)mr * R + r];
        if (counts) {  // ScaledTo workload.go:258-275
          int nc = counts[mr - ps0];
          if (count != 0 && count != nc) {
            for (int r = 0; r < R; r++) if (mk & (1u << r)) mreq[m][r] = mreq[m][r] / count * nc;
            count = nc;
          }
        }
        if (coversPods) { mreq[m][s.pods_resource] = count; mk |= 1u << s.pods_resource; }  // :585-587
        mmask[m] = mk; mcount[m] = count;
        for (int r = 0; r < R; r++) if (mk & (1u << r)) reqs[r] += mreq[m][r];  // requests.Add :630
        mask |= mk;
        okMask &= s.ps_flavor_ok[mr];  // checkFlavorForPodSets walks every podset of the group :915-941
      }
      PodSetAssign grp;
      bool hasReasons = false, failed = false;
      for (int r = 0; r < R; r++) {  // :639-661
        if (!(mask & (1u << r))) continue;
        if (reqs[r] == 0 && rgByResource(cq, r) < 0) continue;
        if (grp.flavor[r] >= 0) continue;
        if (!findFlavorForPodSet(wl, row, okMask, p, reqs, mask, r, a.usage, useLast, grp, hasReasons)) { failed = true; break; }
      }
      if (failed) hasReasons = true;
      for (int m = 0; m < nm; m++) {  // :664-675
        PodSetAssign psa; psa.count = mcount[m];
        if (!failed)
          for (int r = 0; r < R; r++)  // FilterKeys(groupFlavors, keys(podSet.Requests)) :666
            if ((mmask[m] & (1u << r)) && grp.flavor[r] >= 0) { psa.flavor[r] = grp.flavor[r]; psa.mode[r] = grp.mode[r]; psa.tried[r] = grp.tried[r]; psa.borrow[r] = grp.borrow[r]; }
        psa.hasReasons = hasReasons;
        psa.nFlavors = 0;
        // Assignment.append :717-738
        for (int r = 0; r < R; r++) {
          if (psa.flavor[r] < 0) continue;
          psa.nFlavors++;
          if (psa.borrow[r] > a.borrowing) a.borrowing = psa.borrow[r];
          a.usage.add(psa.flavor[r] * R + r, mreq[m][r]);
        }
        a.ps.push_back(psa);
      }
      if (failed) return a;  // :677-679
    }
    return a;
  }
  // TotalRequestsFor :198-218
  UsageVec totalRequestsFor(int wl, const Assignment &a) {
    UsageVec u;
    int ps0 = s.wl_ps_start[wl];
    int cq = s.wl_cq[wl];
    bool coversPods = s.pods_resource >= 0 && rgByResource(cq, s.pods_resource) >= 0;
    for (size_t i = 0; i < a.ps.size(); i++) {
      int row = ps0 + (int)i;
      int count = s.ps_count[row], nc = a.ps[i].count;
      for (int r = 0; r < R; r++) {
        i64 q;
        if (coversPods && r == s.pods_resource) q = count;  // Requests[pods] was set to Count before scaling (:585-587)
        else if (!(s.ps_req_mask[row] & (1u << r))) continue;
        else q = s.ps_req[(size_t)row * R + r];
        if (count != 0 && count != nc) q = q / count * nc;
        if (q == 0) continue;
        int f = a.ps[i].flavor[r];
        if (f < 0) continue;
        u.add(f * R + r, q);
      }
    }
    return u;
  }
  std::vector<Target> getTargetsFor(int wl, const Assignment &a) {  // GetTargets preemption.go:127-146
    PCtx c; c.p = {s.wl_cq[wl], s.wl_priority[wl], s.wl_ts[wl]};
    for (auto &ps : a.ps)  // flavorResourcesNeedPreemption :480-490
      for (int r = 0; r < R; r++)
        if (ps.flavor[r] >= 0 && ps.mode[r] == KB_MODE_PREEMPT) {
          int fr = ps.flavor[r] * R + r;
          if (std::find(c.frsNeed.begin(), c.frsNeed.end(), fr) == c.frsNeed.end()) c.frsNeed.push_back(fr);
        }
    c.workloadUsage = totalRequestsFor(wl, a);
    return getTargets(c);
  }

  // ---- scheduler.go ----------------------------------------------------------
  bool canBePartiallyAdmitted(int wl) {  // workload.go:514-522
    for (int row = s.wl_ps_start[wl]; row < s.wl_ps_start[wl + 1]; row++)
      if (s.ps_min_count[row] >= 0 && s.ps_count[row] > s.ps_min_count[row]) return true;
    return false;
  }
  void getAssignments(Entry &e) {  // getInitialAssignments :584-625
    int wl = e.wl;
    Assignment full = assign(wl, nullptr);
    int arm = full.repMode();
    if (arm == KB_MODE_FIT) { e.a = full; return; }
    if (arm == KB_MODE_PREEMPT) {
      std::vector<Target> t = getTargetsFor(wl, full);
      if (!t.empty()) { e.a = full; e.targets = t; return; }
    }
    if ((s.flags & KB_F_PARTIAL_ADMISSION) && canBePartiallyAdmitted(wl)) {
      // PodSetReducer podset_reducer.go:37-86
      int ps0 = s.wl_ps_start[wl], np = s.wl_ps_start[wl + 1] - ps0;
      std::vector<int32_t> fullCounts(np), deltas(np), cur(np);
      int total = 0;
      for (int i = 0; i < np; i++) {
        fullCounts[i] = s.ps_count[ps0 + i];
        int mc = s.ps_min_count[ps0 + i] >= 0 ? s.ps_min_count[ps0 + i] : fullCounts[i];
        deltas[i] = fullCounts[i] - mc; total += deltas[i];
      }
      if (total > 0) {
        int lastGood = 0; bool haveGood = false; Assignment goodA; std::vector<Target> goodT;
        auto fitsFn = [&](int i) -> bool {
          for (int k = 0; k < np; k++) cur[k] = fullCounts[k] - (int32_t)((i64)deltas[k] * i / total);
          Assignment as = assign(wl, cur.data());
          int m = as.repMode();
          if (m == KB_MODE_FIT) { lastGood = i; haveGood = true; goodA = as; goodT.clear(); return true; }
          if (m == KB_MODE_PREEMPT) {
            std::vector<Target> t = getTargetsFor(wl, as);
            if (!t.empty()) { lastGood = i; haveGood = true; goodA = as; goodT = t; return true; }
          }
          return false;
        };
        // sort.Search(total+1, f): smallest i in [0,total+1) with f(i) true
        int lo = 0, hi = total + 1;
        while (lo < hi) { int mid = lo + (hi - lo) / 2; if (!fitsFn(mid)) lo = mid + 1; else hi = mid; }
        if (haveGood && lo == lastGood) { e.a = goodA; e.targets = goodT; return; }
      }
    }
    e.a = full;
  }
  // quotaResourcesToReserve :530-548
  UsageVec resourcesToReserve(Entry &e, int cq) {
    if (e.a.repMode() != KB_MODE_PREEMPT) return e.a.usage;
    UsageVec r;
    for (auto &p : e.a.usage.v) {
      int fr = p.first; i64 u = p.second;
      i64 nominal = Nominal(cq, fr), bl = s.borrow_limit[(size_t)cq * FR + fr];
      if (e.a.borrowing > 0) {
        if (bl == KB_NO_LIMIT) r.add(fr, u);
        else r.add(fr, std::min(u, nominal + bl - U(cq, fr)));
      } else r.add(fr, std::max<i64>(0, std::min(u, nominal - U(cq, fr))));
    }
    return r;
  }
  // entryComparer.less fair_sharing_iterator.go:166-199
  static bool fsLessRaw(uint32_t flags, int aBorrowing, int aPrio, i64 aTs, const DRS &da, int bBorrowing, int bPrio, i64 bTs, const DRS &db) {
    if (flags & KB_F_FS_PRIORITIZE_NON_BORROWING) {
      bool ab = aBorrowing > 0, bb = bBorrowing > 0;
      if (ab != bb) return !ab;
    }
    int c = compareDRS(da, db);
    if (c != 0) return c == -1;
    if (flags & KB_F_PRIORITY_SORTING_WITHIN_COHORT) {
      if (aPrio != bPrio) return aPrio > bPrio;
    }
    return aTs < bTs;
  }
  bool fsLess(const Entry &a, const Entry &b, const DRS &da, const DRS &db) {
    return fsLessRaw(s.flags, a.a.borrowing, s.wl_priority[a.wl], s.wl_ts[a.wl], da, b.a.borrowing, s.wl_priority[b.wl], s.wl_ts[b.wl], db);
  }
  // runTournament :120-153 ; drs[node][entry] looked up by (parent cohort, entry)
  int runTournament(int cohort, std::vector<Entry> &entries, const std::vector<int> &cqToEntry,
                    const std::vector<std::vector<std::pair<int, DRS>>> &drsByCohort) {
    std::vector<int> cands;
    for (int ch : childCohorts[cohort]) { int c = runTournament(ch, entries, cqToEntry, drsByCohort); if (c >= 0) cands.push_back(c); }
    for (int cq : childCqs[cohort]) if (cqToEntry[cq] >= 0) cands.push_back(cqToEntry[cq]);
    if (cands.empty()) return -1;
    auto drsOf = [&](int e) -> DRS { for (auto &p : drsByCohort[cohort]) if (p.first == e) return p.second; return DRS(); };
    int best = cands[0];
    for (size_t i = 1; i < cands.size(); i++) if (fsLess(entries[cands[i]], entries[best], drsOf(cands[i]), drsOf(best))) best = cands[i];
    return best;
  }

  void schedule(kb_cycle_out *out) {  // schedule :218-427
    int H = s.n_heads;
    std::vector<Entry> entries(H);
    for (int i = 0; i < H; i++) { entries[i].wl = s.heads[i]; getAssignments(entries[i]); }  // nominate :464-501

    // iterator order
    std::vector<int> order;
    order.reserve(H);
    if (!fair) {  // makeClassicalIterator :778-817
      for (int i = 0; i < H; i++) order.push_back(i);
      bool prio = s.flags & KB_F_PRIORITY_SORTING_WITHIN_COHORT;
      std::sort(order.begin(), order.end(), [&](int x, int y) {
        const Entry &a = entries[x], &b = entries[y];
        bool aq = s.wl_has_quota_reservation && s.wl_has_quota_reservation[a.wl];  // scheduler.go:781-789
        bool bq = s.wl_has_quota_reservation && s.wl_has_quota_reservation[b.wl];
        if (aq != bq) return aq;
        if (a.a.borrowing != b.a.borrowing) return a.a.borrowing < b.a.borrowing;
        if (prio && s.wl_priority[a.wl] != s.wl_priority[b.wl]) return s.wl_priority[a.wl] > s.wl_priority[b.wl];
        if (s.wl_ts[a.wl] != s.wl_ts[b.wl]) return s.wl_ts[a.wl] < s.wl_ts[b.wl];
        return x < y;  // canonical tie-break: position in the cycle's entry list
      });
    }
    std::vector<char> preempted(s.n_adm, 0);  // PreemptedWorkloads
    std::vector<int> preemptedList;
    std::vector<int> rootRank(N, 0);
    auto process = [&](int ei) {  // loop body :269-401
      Entry &e = entries[ei];
      int cq = s.wl_cq[e.wl];
      e.rank = rootRank[rootOf(cq)]++;
      int mode = e.a.repMode();
      if (mode == KB_MODE_NOFIT) { e.decision = KB_DEC_NOFIT; return; }
      if (mode == KB_MODE_PREEMPT && e.targets.empty()) {
        e.decision = KB_DEC_PREEMPT_NO_TARGETS;
        if (s.cq_reclaim_within[cq] != KB_POLICY_ANY) addUsageVec(cq, resourcesToReserve(e, cq));  // CanAlwaysReclaim policy.go:27-29
        return;
      }
      for (auto &t : e.targets) if (preempted[t.adm]) { e.decision = KB_DEC_SKIPPED_OVERLAP; return; }
      // fits :503-511
      for (int a : preemptedList) removeAdm(a);
      for (auto &t : e.targets) removeAdm(t.adm);
      bool ok = fitsVec(cq, e.a.usage);
      for (int a : preemptedList) addAdm(a);
      for (auto &t : e.targets) addAdm(t.adm);
      if (!ok) { e.decision = KB_DEC_SKIPPED_NO_FIT; return; }
      for (auto &t : e.targets) { preempted[t.adm] = 1; preemptedList.push_back(t.adm); }
      addUsageVec(cq, e.a.usage);
      e.decision = (mode == KB_MODE_PREEMPT) ? KB_DEC_PREEMPTING : KB_DEC_ASSUMED;
    };
    if (!fair) {
      for (int ei : order) process(ei);
    } else {  // fairSharingIterator fair_sharing_iterator.go:36-118
      std::vector<int> cqToEntry(Q, -1);
      for (int i = 0; i < H; i++) cqToEntry[s.wl_cq[entries[i].wl]] = i;  // later entries overwrite (:52-54)
      for (int cq = 0; cq < Q; cq++) if (cqToEntry[cq] >= 0 && !hasParent(cq)) { int e = cqToEntry[cq]; cqToEntry[cq] = -1; process(e); }
      for (int root = Q; root < N; root++) {
        if (hasParent(root)) continue;
        std::vector<int> cqs; subtreeClusterQueues(root, cqs);
        while (true) {
          bool any = false;
          for (int cq : cqs) if (cqToEntry[cq] >= 0) { any = true; break; }
          if (!any) break;
          // computeDRS :206-229
          std::vector<std::vector<std::pair<int, DRS>>> drsByCohort(N);
          for (int cq : cqs) {
            int ei = cqToEntry[cq];
            if (ei < 0) continue;
            addUsageVec(cq, entries[ei].a.usage);  // assignmentUsage (netUsage :519-528; no quota reservation)
            DRS d = dominantResourceShare(cq);
            for (int anc = s.parent[cq]; anc >= 0; anc = s.parent[anc]) {
              drsByCohort[anc].push_back({ei, d});
              d = dominantResourceShare(anc);
            }
            removeUsageVec(cq, entries[ei].a.usage);
          }
          int win = runTournament(root, entries, cqToEntry, drsByCohort);
          cqToEntry[s.wl_cq[entries[win].wl]] = -1;
          process(win);
        }
      }
    }
    writeOut(entries, out);
  }

  void writeOut(std::vector<Entry> &entries, kb_cycle_out *out) {
    int H = s.n_heads;
    int nt = 0;
    for (int i = 0; i < H; i++) {
      Entry &e = entries[i];
      out->decision[i] = (uint8_t)e.decision;
      out->mode[i] = (uint8_t)e.a.repMode();
      out->borrow[i] = e.a.borrowing;
      out->commit_rank[i] = e.rank;
      int np = s.wl_ps_start[e.wl + 1] - s.wl_ps_start[e.wl];
      for (int k = 0; k < np; k++) {
        size_t row = (size_t)s.wl_ps_start[e.wl] + k;
        for (int r = 0; r < R; r++) { out->ps_flavor[row * R + r] = -1; out->ps_res_mode[row * R + r] = -1; out->ps_tried_idx[row * R + r] = -1; }
        out->ps_count[row] = s.ps_count[row];
        if (k < (int)e.a.ps.size()) {
          const PodSetAssign &p = e.a.ps[k];
          out->ps_count[row] = p.count;
          for (int r = 0; r < R; r++) {
            out->ps_flavor[row * R + r] = p.flavor[r];
            out->ps_res_mode[row * R + r] = p.flavor[r] >= 0 ? p.mode[r] : (int8_t)-1;
            out->ps_tried_idx[row * R + r] = p.flavor[r] >= 0 ? p.tried[r] : (int8_t)-1;
          }
        }
      }
      out->tgt_start[i] = nt;
      for (auto &t : e.targets) {
        if (nt < out->tgt_capacity) { out->tgt_adm[nt] = t.adm; out->tgt_reason[nt] = (uint8_t)t.reason; }
        nt++;
      }
    }
    out->tgt_start[H] = nt;
    out->n_targets = nt;
    if (out->node_usage) memcpy(out->node_usage, usage.data(), sizeof(i64) * (size_t)N * FR);
  }
};

}  // namespace

extern "C" {

// K1 parity surface: SubtreeQuota / Usage / Available / PotentialAvailable / DRS.
// wl_req: optional [F*R] request added to every node's usage when computing DRS
// (the wlReq argument of dominantResourceShare, fair_sharing.go:126).
int32_t ko_tree_eval_req(const kb_snapshot *s, const int64_t *wl_req, kb_tree_out *out) {
  Oracle o(*s);
  size_t n = (size_t)o.N * o.FR;
  if (out->subtree_quota) memcpy(out->subtree_quota, o.subtree.data(), n * sizeof(i64));
  if (out->usage) memcpy(out->usage, o.usage.data(), n * sizeof(i64));
  for (int q = 0; q < o.Q; q++)
    for (int fr = 0; fr < o.FR; fr++) {
      if (out->available) out->available[(size_t)q * o.FR + fr] = o.Available(q, fr);
      if (out->potential_available) out->potential_available[(size_t)q * o.FR + fr] = o.potentialAvailable(q, fr);
    }
  for (int nd = 0; nd < o.N; nd++) {
    DRS d = o.dominantResourceShare(nd, wl_req);
    if (out->drs_rounded) out->drs_rounded[nd] = Oracle::roundedWeightedShare(d);
    if (out->drs_resource) out->drs_resource[nd] = d.dominantResource;
    if (out->drs_borrowing) out->drs_borrowing[nd] = d.borrowing;
  }
  return 0;
}

int32_t ko_tree_eval(const kb_snapshot *s, kb_tree_out *out) { return ko_tree_eval_req(s, nullptr, out); }

int32_t ko_run_cycle(const kb_snapshot *s, kb_cycle_out *out) {
  Oracle o(*s);
  o.schedule(out);
  return out->n_targets > out->tgt_capacity ? KB_ERR_CAPACITY : 0;
}

// flavorassigner.Assign(nil) for workload `wl` with the reference's stub preemption oracle
// (TestAssignFlavors flavorassigner_test.go:165): stub_mode[fr] in {1 NoCandidates, 2 Preempt,
// 3 Reclaim} or -1 for the default (Preempt, 0).  Outputs like one entry of ko_run_cycle.
int32_t ko_assign_stub(const kb_snapshot *s, int32_t wl, const int8_t *stub_mode, const int32_t *stub_borrow,
                       int8_t *ps_flavor, int8_t *ps_res_mode, int8_t *ps_tried, int32_t *borrowing, int64_t *usage_fr) {
  Oracle o(*s);
  o.useStub = true; o.stubMode = stub_mode; o.stubBorrow = stub_borrow;
  Assignment a = o.assign(wl, nullptr);
  int np = s->wl_ps_start[wl + 1] - s->wl_ps_start[wl];
  for (int k = 0; k < np; k++)
    for (int r = 0; r < o.R; r++) {
      bool have = k < (int)a.ps.size() && a.ps[k].flavor[r] >= 0;
      ps_flavor[(size_t)k * o.R + r] = have ? a.ps[k].flavor[r] : (int8_t)-1;
      ps_res_mode[(size_t)k * o.R + r] = have ? a.ps[k].mode[r] : (int8_t)-1;
      ps_tried[(size_t)k * o.R + r] = have ? a.ps[k].tried[r] : (int8_t)-1;
    }
  *borrowing = a.borrowing;
  for (int fr = 0; fr < o.FR; fr++) usage_fr[fr] = -1;
  for (auto &p : a.usage.v) usage_fr[p.first] = p.second;
  return a.repMode();
}

// Single preemption query (TestPreemption-style goldens): targets for workload
// `wl` given an explicit assignment (flavor per podset/resource + mode).
int32_t ko_get_targets(const kb_snapshot *s, int32_t wl, const int8_t *ps_flavor, const int8_t *ps_res_mode,
                       const int32_t *ps_assigned_count, int32_t *tgt_adm, uint8_t *tgt_reason, int32_t cap) {
  Oracle o(*s);
  Assignment a;
  int ps0 = s->wl_ps_start[wl], np = s->wl_ps_start[wl + 1] - ps0;
  for (int k = 0; k < np; k++) {
    PodSetAssign p; p.count = ps_assigned_count ? ps_assigned_count[k] : s->ps_count[ps0 + k]; p.hasReasons = true;
    for (int r = 0; r < o.R; r++) {
      p.flavor[r] = ps_flavor[(size_t)k * o.R + r]; p.mode[r] = ps_res_mode[(size_t)k * o.R + r];
      if (p.flavor[r] >= 0) p.nFlavors++;
    }
    a.ps.push_back(p);
  }
  std::vector<Target> t = o.getTargetsFor(wl, a);
  int n = 0;
  for (auto &x : t) { if (n < cap) { tgt_adm[n] = x.adm; tgt_reason[n] = (uint8_t)x.reason; } n++; }
  return n;
}


// resourcesToReserve (scheduler.go:530-548) for one ClusterQueue and an explicit assignment, the surface
// TestResourcesToReserve (scheduler_test.go:9705) checks.  usage / out: [F*R], -1 = cell absent.
int32_t ko_resources_to_reserve(const kb_snapshot *s, int32_t cq, int32_t mode, int32_t borrowing, const int64_t *usage, int64_t *out) {
  Oracle o(*s);
  Entry e;
  e.wl = -1;
  PodSetAssign p;
  p.count = 1; p.hasReasons = mode != KB_MODE_FIT; p.nFlavors = 1; p.flavor[0] = 0; p.mode[0] = (int8_t)mode;
  e.a.ps.push_back(p);
  e.a.borrowing = borrowing;
  for (int fr = 0; fr < o.FR; fr++) { out[fr] = -1; if (usage[fr] >= 0) e.a.usage.add(fr, usage[fr]); }
  UsageVec r = o.resourcesToReserve(e, cq);
  for (auto &c : r.v) out[c.first] = c.second;
  return 0;
}

// entryComparer.less (fair_sharing_iterator.go:166-199) on explicit operands, the surface TestEntryComparerLess
// (scheduler_test.go:9577) checks.  ratio / weight are DRS.unweightedRatio / fairWeight.
int32_t ko_entry_less(uint32_t flags, int32_t a_borrowing, int32_t a_prio, int64_t a_ts, double a_ratio, double a_weight,
                      int32_t b_borrowing, int32_t b_prio, int64_t b_ts, double b_ratio, double b_weight) {
  DRS da, db;
  da.unweightedRatio = a_ratio; da.fairWeight = a_weight; db.unweightedRatio = b_ratio; db.fairWeight = b_weight;
  return Oracle::fsLessRaw(flags, a_borrowing, a_prio, a_ts, da, b_borrowing, b_prio, b_ts, db) ? 1 : 0;
}

// SatisfiesPreemptionPolicy (preemption/common/preemption_policy.go:30-48) of preemptor (priority, timestamp) against
// admitted workload `adm` — TestSatisfiesPreemptionPolicy (preemption_policy_test.go:32).
int32_t ko_satisfies_policy(const kb_snapshot *s, int32_t pre_prio, int64_t pre_ts, int32_t adm, int32_t policy) {
  Oracle o(*s);
  Preemptor p{0, pre_prio, pre_ts};
  return o.satisfiesPreemptionPolicy(p, adm, policy) ? 1 : 0;
}

// Every admitted workload of the snapshot sorted by CandidatesOrdering (preemption/common/ordering.go:41-100) for
// a preemptor in ClusterQueue `cq` — TestCandidatesOrdering (preemption_test.go:4525).
int32_t ko_sort_candidates(const kb_snapshot *s, int32_t cq, int32_t *order) {
  Oracle o(*s);
  std::vector<int> v(s->n_adm);
  for (int i = 0; i < s->n_adm; i++) v[i] = i;
  std::stable_sort(v.begin(), v.end(), [&](int a, int b) { return o.candidatesOrdering(a, b, cq) < 0; });
  for (int i = 0; i < s->n_adm; i++) order[i] = v[i];
  return 0;
}

// isPreferred (flavorassigner.go:410-441) on explicit granular modes — TestIsPreferred (flavorassigner_test.go:3885).
// pm: 0 noFit, 1 noCandidates, 2 preempt, 3 reclaim, 4 fit; pref: KB_PREF_*.
int32_t ko_is_preferred(int32_t a_pm, int32_t a_borrow, int32_t b_pm, int32_t b_borrow, int32_t pref) {
  kb_snapshot z{};
  Oracle o(z);
  return o.isPreferred(GranularMode{a_pm, a_borrow}, GranularMode{b_pm, b_borrow}, pref) ? 1 : 0;
}
}  // extern "C"
