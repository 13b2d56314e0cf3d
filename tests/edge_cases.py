"""Edge-case snapshots shared by the oracle (CPU) and device (GPU) tests: empty and ragged inputs."""
import numpy as np

from kueue_b200 import abi
from kueue_b200.api import MakeAdmission, MakeClusterQueue, MakeCohort, MakeFlavorQuotas, MakePodSet, MakeWorkload, flatten


def _cq(name, cohort=None, nominal="4"):
    c = MakeClusterQueue(name).ResourceGroup(MakeFlavorQuotas("default").Resource("cpu", nominal))
    return c.Cohort(cohort) if cohort else c


def _wl(name, cq, cpu="1", count=1, prio=0):
    return MakeWorkload(name).ClusterQueue(cq).Priority(prio).PodSets(MakePodSet("main", count).Request("cpu", cpu))


def edge_snapshots():
    out = {}
    # no heads at all: the cycle has nothing to decide
    out["no pending workloads"] = flatten([_cq("a"), _cq("b", "co")])[0]
    # pending workloads exist but none is a head this cycle
    s, _ = flatten([_cq("a")], pending=[_wl("w", "a")], heads=[])
    out["pending but no heads"] = s
    # one ClusterQueue, one workload
    out["single workload"] = flatten([_cq("a")], pending=[_wl("w", "a")])[0]
    # a cohort without ClusterQueues next to a populated one; a ClusterQueue without quota
    out["empty cohort and zero-quota queue"] = flatten([_cq("a", "co"), _cq("z", "co", "0")], [MakeCohort("co"), MakeCohort("lonely")],
                                                        pending=[_wl("w1", "a"), _wl("w2", "z", "3")])[0]
    # request for a resource the ClusterQueue does not cover, and one that exceeds every quota
    out["uncovered resource and oversize"] = flatten(
        [_cq("a")], pending=[MakeWorkload("gpu").ClusterQueue("a").PodSets(MakePodSet("main", 1).Request("example.com/gpu", "1")),
                             ], extra_resources=["example.com/gpu"])[0]
    out["oversize request"] = flatten([_cq("a")], pending=[_wl("big", "a", "100")])[0]
    # ragged podsets: 1, 3 and 8 podsets in one cycle
    pend = [_wl("p1", "a")]
    pend.append(MakeWorkload("p3").ClusterQueue("b").PodSets(*[MakePodSet(f"s{i}", i + 1).Request("cpu", "100m") for i in range(3)]))
    pend.append(MakeWorkload("p8").ClusterQueue("c").PodSets(*[MakePodSet(f"s{i}", 1).Request("cpu", "100m") for i in range(8)]))
    out["ragged podsets"] = flatten([_cq("a", "co"), _cq("b", "co"), _cq("c", "co")], pending=pend)[0]
    # every admitted workload is a preemption candidate of a single head; lone ClusterQueue, within-queue preemption
    cq = _cq("a", nominal="4").Preemption("LowerPriority", "Never")
    adm = [MakeWorkload(f"low{i}").Priority(-1).Request("cpu", "1").ReserveQuota(MakeAdmission("a").Assignment("cpu", "default", "1"), 1000 + i) for i in range(4)]
    out["lone queue preempts all"] = flatten([cq], pending=[_wl("hi", "a", "4", prio=10)], admitted=adm, now_ns=5000)[0]
    # admitted workloads but nothing pending
    out["admitted only"] = flatten([cq], admitted=adm, now_ns=5000)[0]
    return out
