"""Builders shared by the CPU (oracle) and GPU (device) replays of tests/golden/schedule_cases.py."""
import numpy as np

from kueue_b200 import abi
from kueue_b200.api import (MakeClusterQueue, MakeCohort, MakeFlavorQuotas, MakePodSet, MakeWorkload, flatten)
from tests.golden.schedule_cases import NOW, S


def entry_ordering_snapshot(tc):
    """One root cohort holding every entry of a TestEntryOrdering case, each in its own ClusterQueue, so the
    classical iterator (scheduler.go:778-817) orders them together and `commit_rank` is the pop order.

    Borrowing level of an entry = height of the lowest subtree its request fits in
    (FindHeightOfLowestSubtreeThatFits, hierarchical_preemption.go:214-227):
      0: ClusterQueue with nominal quota under cohort m1;
      1: ClusterQueue without quota under m1, which holds a lender ClusterQueue      -> fits at m1   (height 1);
      2: ClusterQueue without quota under m2 (empty), lender directly under the root -> fits at root (height 2).
    """
    cohorts = [MakeCohort("root"), MakeCohort("m1").Parent("root"), MakeCohort("m2").Parent("root")]

    def cq(name, cohort, nominal):
        return MakeClusterQueue(name).Cohort(cohort).ResourceGroup(MakeFlavorQuotas("default").Resource("cpu", str(nominal)))

    cqs = [cq("lender-m1", "m1", 1000), cq("lender-root", "root", 1000)]
    pending = []
    for name, created, prio, borrowing, cond in tc["input"]:
        cqs.append(cq("cq-" + name, "m2" if borrowing == 2 else "m1", 10 if borrowing == 0 else 0))
        w = MakeWorkload(name).ClusterQueue("cq-" + name).Priority(prio).Creation(NOW + created * S).PodSets(MakePodSet("main", 1).Request("cpu", "1"))
        if cond:
            w.Condition(cond[0], cond[1], cond[2], NOW + cond[3] * S)
        pending.append(w)
    flags = abi.FLAGS_DEFAULT
    if not tc["priority_sorting"]:
        flags &= ~abi.F_PRIORITY_SORTING_WITHIN_COHORT
    snap, idx = flatten(cqs, cohorts, pending=pending, flags=flags, pods_ready_requeuing=tc["ordering"], now_ns=NOW + 100 * S)
    want_borrow = np.array([b for _, _, _, b, _ in tc["input"]], np.int32)
    return snap, idx, want_borrow


def order_of(out, tc):
    names = [n for n, *_ in tc["input"]]
    return [names[i] for i in np.argsort(out.commit_rank, kind="stable")]


def reducer_snapshot(tc):
    """TestSearch's predicate `total pods <= countLimit` as a quota: every pod asks for one unit of a resource whose
    nominal quota is countLimit, so PodSetReducer.Search (podset_reducer.go:56-86) runs inside the real cycle."""
    cq = MakeClusterQueue("cq").ResourceGroup(MakeFlavorQuotas("default").Resource("example.com/slot", str(tc["limit"])))
    pss = []
    for i, (count, min_count) in enumerate(tc["podsets"]):
        p = MakePodSet(f"ps{i + 1}", count).Request("example.com/slot", "1")
        if min_count is not None:
            p.SetMinimumCount(min_count)
        pss.append(p)
    w = MakeWorkload("wl").ClusterQueue("cq").PodSets(*pss)
    return flatten([cq], pending=[w], now_ns=NOW)


def lending_snapshot(remaining):
    from kueue_b200.api import MakeAdmission
    from tests.golden.schedule_cases import LENDING_WORKLOADS
    cqs = [MakeClusterQueue("lend-a").Cohort("lend").ResourceGroup(MakeFlavorQuotas("default").Resource("cpu", "10", "", "4")).Preemption("LowerPriority", "LowerPriority"),
           MakeClusterQueue("lend-b").Cohort("lend").ResourceGroup(MakeFlavorQuotas("default").Resource("cpu", "10", "", "6")).Preemption("Never", "Any")]
    adm = [MakeWorkload(n).Request("cpu", str(LENDING_WORKLOADS[n][1])).ReserveQuota(MakeAdmission(LENDING_WORKLOADS[n][0]).Assignment("cpu", "default", str(LENDING_WORKLOADS[n][1])), NOW)
           for n in remaining]
    return flatten(cqs, admitted=adm, now_ns=NOW)
