// kueue_oracle_tas.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Single-threaded CPU restatement of the topology-aware placement of one TAS ResourceFlavor
// (pkg/cache/scheduler/tas_flavor_snapshot.go), the parity checker of kb_tas_find.  Nothing under kueue_b200/ may
// link, import or call this file.  Every function cites the reference file:line it restates (paths relative to
// /root/reference).  Scope = the scope of kb_tas_find (include/kueue_b200.h): BestFit / LeastFreeCapacity placement
// of podsets without leader/worker groups, balanced placement, multi-layer slices, node replacement or elastic slices.
//
// Pinning (parity pinned): TestFindTopologyAssignments (pkg/cache/scheduler/tas_cache_test.go:55), transcribed by
// tools/transcribe_tas.py into tests/golden/tas_cases.json — tests/test_oracle_golden_tas.py.

#include "../include/kueue_b200.h"

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

typedef int64_t i64;

struct Domain {  // tas_flavor_snapshot.go:62-103 (leader fields dropped: no leader podsets in scope)
  int32_t state = 0, sliceState = 0;
  int parent = -1;
  std::vector<int> children;
};

struct Tas {
  const kb_tas_topology &t;
  const kb_tas_requests &r;
  int L, nleaf, leaf0, R;
  std::vector<Domain> dom;
  std::vector<i64> assumed;          // [nleaf][R] assumedUsage of the current chain
  std::vector<uint32_t> assumedMask; // [nleaf]

  Tas(const kb_tas_topology &tt, const kb_tas_requests &rr) : t(tt), r(rr) {
    L = t.n_levels; R = t.n_resource;
    leaf0 = t.level_start[L - 1]; nleaf = t.n_domains - leaf0;
    dom.resize(t.n_domains);
    for (int d = 0; d < t.n_domains; d++) { dom[d].parent = t.parent[d]; if (t.parent[d] >= 0) dom[t.parent[d]].children.push_back(d); }
    assumed.assign((size_t)nleaf * R, 0); assumedMask.assign(nleaf, 0);
  }
  int levelOf(int d) const { int l = 0; while (d >= t.level_start[l + 1]) l++; return l; }

  // Requests.CountIn requests.go:172-205 for requests = SinglePodRequests + {pods: 1}
  int32_t countIn(int q, const i64 *cap, uint32_t capMask) const {
    bool have = false; int32_t result = 0;
    uint32_t keys = r.request_mask[q] | (1u << t.pods_resource);
    for (int k = 0; k < R; k++) {
      if (!((keys >> k) & 1)) continue;
      i64 v = k == t.pods_resource ? 1 : r.pod_request[(size_t)q * R + k];
      if (!((capMask >> k) & 1) && v != 0) return 0;
      int32_t count = v == 0 ? INT32_MAX : (int32_t)(cap[k] / v);
      if (!have || count < result) { result = count; have = true; }
    }
    return have ? result : 0;
  }

  // fillInCounts :1517-1608 + fillInCountsHelper :1620-1672
  void fillInCounts(int q, int sliceSize, int sliceLevel, bool simulateEmpty) {
    for (auto &d : dom) { d.state = 0; d.sliceState = 0; }
    std::vector<i64> rem(R);
    for (int lf = 0; lf < nleaf; lf++) {
      if (r.leaf_ok && !((r.leaf_ok[(size_t)q * ((nleaf + 31) / 32) + lf / 32] >> (lf % 32)) & 1)) continue;  // taints / selectors :1541-1571
      uint32_t mask = t.cap_mask[lf] | assumedMask[lf];
      for (int k = 0; k < R; k++) rem[k] = t.free_capacity[(size_t)lf * R + k] - assumed[(size_t)lf * R + k];
      if (!simulateEmpty) {
        mask |= t.usage_mask[lf];
        for (int k = 0; k < R; k++) rem[k] -= t.tas_usage[(size_t)lf * R + k];
      }
      dom[leaf0 + lf].state = countIn(q, rem.data(), mask);
    }
    for (int l = L - 1; l >= 0; l--)
      for (int d = t.level_start[l]; d < t.level_start[l + 1]; d++) {
        Domain &D = dom[d];
        if (!D.children.empty()) {
          int32_t cap = 0, sl = 0;
          for (int c : D.children) { cap += dom[c].state; sl += dom[c].sliceState; }
          D.state = cap; D.sliceState = sl;
        }
        if (l == sliceLevel) D.sliceState = D.state / sliceSize;
      }
  }

  bool leastFree(bool unconstrained, uint32_t flags) const { return unconstrained && (flags & KB_TAS_PROFILE_MIXED); }  // :1291-1294
  // sortedDomains :1495-1515 (levelValues order = domain index inside a level)
  std::vector<int> sorted(std::vector<int> v, bool lfc) const {
    std::sort(v.begin(), v.end(), [&](int a, int b) {
      if (dom[a].sliceState != dom[b].sliceState) return lfc ? dom[a].sliceState < dom[b].sliceState : dom[a].sliceState > dom[b].sliceState;
      if (dom[a].state != dom[b].state) return dom[a].state < dom[b].state;
      return a < b;
    });
    return v;
  }
  int bestFitBy(const std::vector<int> &v, size_t from, int32_t needed, bool slices) const {  // findBestFitDomainBy :1183-1198
    auto st = [&](int d) { return slices ? dom[d].sliceState : dom[d].state; };
    int best = v[from]; int32_t bs = st(best);
    for (size_t i = from; i < v.size(); i++) { int32_t s = st(v[i]); if (s >= needed && s < bs) { best = v[i]; bs = s; } }
    return best;
  }

  // findLevelWithFitDomains :1200-1282 (leaderPodSetSize = 0)
  bool findLevel(int levelIdx, bool required, int32_t count, int32_t sliceSize, bool unconstrained, uint32_t flags, int *fitLevel, std::vector<int> *fit) {
    std::vector<int> level;
    for (int d = t.level_start[levelIdx]; d < t.level_start[levelIdx + 1]; d++) level.push_back(d);
    if (level.empty()) return false;
    const bool lfc = leastFree(unconstrained, flags);
    std::vector<int> sd = sorted(level, lfc);
    int top = sd[0];
    int32_t sliceCount = count / sliceSize;
    if (!lfc && dom[top].sliceState >= sliceCount) top = bestFitBy(sd, 0, sliceCount, true);
    if (lfc) {
      for (int c : sd) if (dom[c].sliceState >= sliceCount) { *fitLevel = levelIdx; *fit = {c}; return true; }
      if (required) return false;
    }
    if (dom[top].sliceState < sliceCount) {
      if (required) return false;
      if (levelIdx > 0 && !unconstrained) return findLevel(levelIdx - 1, required, count, sliceSize, unconstrained, flags, fitLevel, fit);
      std::vector<int> res;
      int32_t remaining = sliceCount;
      for (size_t i = 0; remaining > 0 && i < sd.size(); i++) {
        int d = sd[i];
        if (!lfc && dom[d].sliceState >= remaining) d = bestFitBy(sd, i, remaining, true);
        res.push_back(d);
        remaining -= dom[d].sliceState;
      }
      if (remaining > 0) return false;
      *fitLevel = levelIdx; *fit = res;
      return true;
    }
    *fitLevel = levelIdx; *fit = {top};
    return true;
  }

  // updateCountsToMinimumGeneric :1361-1428 (leaderCount = 0); mutates the domains' state / sliceState
  bool updateCounts(const std::vector<int> &domains, int32_t count, int32_t sliceSize, bool lfc, bool slices, std::vector<int> *out) {
    out->clear();
    int32_t remaining = slices ? count / sliceSize : count;
    for (size_t i = 0; i < domains.size(); i++) {
      int d = domains[i];
      if (slices) {
        if (!lfc && dom[d].sliceState >= remaining) d = bestFitBy(domains, i, remaining, true);
        if (dom[d].sliceState >= remaining) { dom[d].state = remaining * sliceSize; dom[d].sliceState = remaining; out->push_back(d); return true; }
        dom[d].state = dom[d].sliceState * sliceSize;
        remaining -= dom[d].sliceState;
        out->push_back(d);
        continue;
      }
      if (!lfc && dom[d].state >= remaining) d = bestFitBy(domains, i, remaining, false);
      if (dom[d].state >= remaining) { dom[d].state = remaining; out->push_back(d); return true; }
      remaining -= dom[d].state;
      out->push_back(d);
    }
    return false;  // errCodeAssumptionsViolated: nil
  }

  // findTopologyAssignment :765-970 for request q; appends (leaf, count) pairs in leaf order
  int place(int q, std::vector<std::pair<int, int>> *asg) {
    const uint32_t flags = r.flags[q];
    const int32_t count = r.count[q], sliceSize = r.slice_size[q];
    const int levelIdx = r.level[q], sliceLevel = r.slice_level[q];
    const bool required = flags & KB_TAS_REQUIRED, unconstrained = flags & KB_TAS_UNCONSTRAINED;
    if (levelIdx < 0 || levelIdx >= L || sliceLevel < 0 || sliceLevel >= L || levelIdx > sliceLevel || sliceSize < 1) return KB_TAS_BAD_REQUEST;
    fillInCounts(q, sliceSize, sliceLevel, flags & KB_TAS_SIMULATE_EMPTY);
    int fitLevel = 0; std::vector<int> cur;
    if (!findLevel(levelIdx, required, count, sliceSize, unconstrained, flags, &fitLevel, &cur)) return KB_TAS_NO_FIT;
    const bool lfc = leastFree(unconstrained, flags);
    std::vector<int> next;
    if (!updateCounts(cur, count, sliceSize, lfc, true, &next)) return KB_TAS_NO_FIT;
    cur = next;
    int lv = fitLevel;
    for (; lv < std::min(L - 1, sliceLevel); lv++) {  // above the slice level: greedy over all children (:901-906)
      std::vector<int> lower;
      for (int d : cur) for (int c : dom[d].children) lower.push_back(c);
      if (!updateCounts(sorted(lower, lfc), count, sliceSize, lfc, true, &next)) return KB_TAS_NO_FIT;
      cur = next;
    }
    for (; lv < L - 1; lv++) {  // at / below the slice level: every parent distributes its own pods (:908-941)
      std::vector<int> all;
      for (int d : cur) {
        if (!updateCounts(sorted(dom[d].children, lfc), dom[d].state, 1, lfc, false, &next)) return KB_TAS_NO_FIT;
        all.insert(all.end(), next.begin(), next.end());
      }
      cur = all;
    }
    std::sort(cur.begin(), cur.end());  // buildAssignment :1455-1466: lexicographic levelValues
    for (int d : cur) if (dom[d].state != 0) asg->push_back({d - leaf0, dom[d].state});  // buildTopologyAssignmentForLevels :1436-1453
    return KB_TAS_OK;
  }
};

}  // namespace

extern "C" int32_t ko_tas_find(const kb_tas_topology *t, const kb_tas_requests *r, kb_tas_out *out) {
  Tas T(*t, *r);
  int n = 0;
  bool chainFailed = false;
  for (int q = 0; q < r->n_req; q++) {
    bool newChain = q == 0 || r->chain[q] != r->chain[q - 1];
    if (newChain) { chainFailed = false; std::fill(T.assumed.begin(), T.assumed.end(), 0); std::fill(T.assumedMask.begin(), T.assumedMask.end(), 0u); }
    out->asg_start[q] = n;
    if (chainFailed) { out->status[q] = -1; continue; }  // FindTopologyAssignmentsForFlavor returns at the first failure (:551-553)
    std::vector<std::pair<int, int>> asg;
    int st = T.place(q, &asg);
    out->status[q] = st;
    if (st != KB_TAS_OK) { chainFailed = true; continue; }
    for (auto &p : asg) {
      if (n < out->capacity) { out->asg_leaf[n] = p.first; out->asg_count[n] = p.second; }
      n++;
      for (int k = 0; k < T.R; k++)  // addAssumedUsage :619-627: SinglePodRequests (no pods) x count
        if ((r->request_mask[q] >> k) & 1) T.assumed[(size_t)p.first * T.R + k] += r->pod_request[(size_t)q * T.R + k] * p.second;
      T.assumedMask[p.first] |= r->request_mask[q];
    }
  }
  out->asg_start[r->n_req] = n;
  out->n_assigned = n;
  return n > out->capacity ? KB_ERR_CAPACITY : KB_OK;
}
