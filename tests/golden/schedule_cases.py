"""Hand transcription of three reference table tests of the cycle's ordering / reservation logic:

  TestGetQueueOrderTimestamp  pkg/workload/workload_test.go:698-778
  TestEntryOrdering           pkg/scheduler/scheduler_test.go:8291-8567
  TestResourcesToReserve      pkg/scheduler/scheduler_test.go:9705-9906

Timestamps are ns relative to an arbitrary `NOW`; only names, priorities, conditions, Borrowing levels and the
expected orders / quantities are taken from the reference.
"""
S = 1_000_000_000
NOW = 1_700_000_000 * S
HOUR = 3600 * S

# ---- TestGetQueueOrderTimestamp: (condition | None) -> {ordering: "creation" | "condition"}
QUEUE_ORDER_TS_CASES = {
    "no condition": (None, {"Eviction": "creation", "Creation": "creation"}),
    "evicted by preemption": (("Evicted", True, "Preempted"), {"Eviction": "creation", "Creation": "creation"}),
    "evicted by PodsReady timeout": (("Evicted", True, "PodsReadyTimeout"), {"Eviction": "condition", "Creation": "creation"}),
    "after eviction": (("Evicted", False, "PodsReadyTimeout"), {"Eviction": "creation", "Creation": "creation"}),
}

# ---- TestEntryOrdering inputs: (name, creation offset s, priority, Borrowing, condition | None)
EVICTED_AT_2 = ("Evicted", True, "PodsReadyTimeout", 2)
ENTRY_INPUT = [
    ("old_borrowing", 0, 0, 1, None),
    ("old", 1, 0, 0, None),
    ("new", 3, 0, 0, None),
    ("high_pri_borrowing", 3, 1, 1, None),
    ("new_high_pri", 4, 1, 0, None),
    ("new_borrowing", 3, 0, 1, None),
    ("evicted_borrowing", 1, 0, 1, EVICTED_AT_2),
    ("recently_evicted", 0, 0, 0, EVICTED_AT_2),
    ("high_pri_borrowing_more", 3, 1, 2, None),
]
ENTRY_INPUT_PREEMPTED = [
    ("old-mid-recently-preempted-in-queue", 0, 1, 0, ("Preempted", True, "InClusterQueue", 5)),
    ("old-mid-recently-reclaimed-while-borrowing", 0, 1, 0, ("Preempted", True, "InCohortReclaimWhileBorrowing", 6)),
    ("old-mid-more-recently-reclaimed-while-borrowing", 0, 1, 0, ("Preempted", True, "InCohortReclaimWhileBorrowing", 7)),
    ("old-mid-not-preempted-yet", 1, 1, 0, None),
    ("preemptor", 7, 2, 0, None),
]
ENTRY_ORDERING_CASES = {
    "Priority sorting is enabled (default) using pods-ready Eviction timestamp (default)": dict(
        input=ENTRY_INPUT, priority_sorting=True, ordering="Eviction",
        want=["new_high_pri", "old", "recently_evicted", "new", "high_pri_borrowing", "old_borrowing", "evicted_borrowing",
              "new_borrowing", "high_pri_borrowing_more"]),
    "Priority sorting is enabled (default) using pods-ready Creation timestamp": dict(
        input=ENTRY_INPUT, priority_sorting=True, ordering="Creation",
        want=["new_high_pri", "recently_evicted", "old", "new", "high_pri_borrowing", "old_borrowing", "evicted_borrowing",
              "new_borrowing", "high_pri_borrowing_more"]),
    "Priority sorting is disabled using pods-ready Eviction timestamp": dict(
        input=ENTRY_INPUT, priority_sorting=False, ordering="Eviction",
        want=["old", "recently_evicted", "new", "new_high_pri", "old_borrowing", "evicted_borrowing", "high_pri_borrowing",
              "new_borrowing", "high_pri_borrowing_more"]),
    "Priority sorting is disabled using pods-ready Creation timestamp": dict(
        input=ENTRY_INPUT, priority_sorting=False, ordering="Creation",
        want=["recently_evicted", "old", "new", "new_high_pri", "old_borrowing", "evicted_borrowing", "high_pri_borrowing",
              "new_borrowing", "high_pri_borrowing_more"]),
    "Some workloads are preempted; Priority sorting is disabled": dict(
        input=ENTRY_INPUT_PREEMPTED, priority_sorting=False, ordering="Eviction",
        want=["old-mid-recently-preempted-in-queue", "old-mid-not-preempted-yet", "old-mid-recently-reclaimed-while-borrowing",
              "preemptor", "old-mid-more-recently-reclaimed-while-borrowing"]),
    "Some workloads are preempted; Priority sorting is enabled": dict(
        input=ENTRY_INPUT_PREEMPTED, priority_sorting=True, ordering="Eviction",
        want=["preemptor", "old-mid-recently-preempted-in-queue", "old-mid-recently-reclaimed-while-borrowing",
              "old-mid-more-recently-reclaimed-while-borrowing", "old-mid-not-preempted-yet"]),
}

# ---- TestResourcesToReserve: ClusterQueue "cq" (cohort "eng"):
#   memory: on-demand nominal 100; spot nominal 0, borrowingLimit 100
#   gpu:    model-a nominal 10, borrowingLimit 0; model-b nominal 10, borrowingLimit 5
MEM, GPU = "memory", "gpu"
CQ_USAGE_A = {("on-demand", MEM): 60, ("spot", MEM): 50, ("model-a", GPU): 6, ("model-b", GPU): 2}
CQ_USAGE_B = {("on-demand", MEM): 60, ("spot", MEM): 50, ("model-a", GPU): 2, ("model-b", GPU): 2}
CQ_USAGE_C = {("on-demand", MEM): 60, ("spot", MEM): 60, ("model-a", GPU): 2, ("model-b", GPU): 10}
RESERVE_CASES = {
    "Reserved memory and gpu less than assignment usage, assignment preempts": dict(
        mode="Preempt", borrowing=0, usage={("on-demand", MEM): 50, ("model-a", GPU): 6}, cq_usage=CQ_USAGE_A,
        want={("on-demand", MEM): 40, ("model-a", GPU): 4}),
    "Reserved memory equal assignment usage, assignment preempts": dict(
        mode="Preempt", borrowing=0, usage={("on-demand", MEM): 30, ("model-a", GPU): 2}, cq_usage=CQ_USAGE_B,
        want={("on-demand", MEM): 30, ("model-a", GPU): 2}),
    "Reserved memory equal assignment usage, assignment fits": dict(
        mode="Fit", borrowing=0, usage={("on-demand", MEM): 50, ("model-a", GPU): 2}, cq_usage=CQ_USAGE_B,
        want={("on-demand", MEM): 50, ("model-a", GPU): 2}),
    "Reserved memory is 0, CQ is borrowing, assignment preempts without borrowing": dict(
        mode="Preempt", borrowing=0, usage={("spot", MEM): 50, ("model-b", GPU): 2}, cq_usage=CQ_USAGE_C,
        want={("spot", MEM): 0, ("model-b", GPU): 0}),
    "Reserved memory cut by nominal+borrowing quota, assignment preempts and borrows": dict(
        mode="Preempt", borrowing=1, usage={("spot", MEM): 50, ("model-b", GPU): 2}, cq_usage=CQ_USAGE_C,
        want={("spot", MEM): 40, ("model-b", GPU): 2}),
    "Reserved memory equal assignment usage, CQ borrowing limit is nil": dict(
        mode="Preempt", borrowing=1, usage={("on-demand", MEM): 50, ("model-b", GPU): 2}, cq_usage=CQ_USAGE_C,
        want={("on-demand", MEM): 50, ("model-b", GPU): 2}),
}

# ---- TestEntryComparerLess scheduler_test.go:9577: (name, creation offset s, Borrowing, DRS (ratio, weight) | None)
#      a missing drsValues entry is Go's zero DRS{} (ratio 0, weight 0); NegativeDRS() = (ratio -1, weight 1)
ENTRY_LESS_CASES = {
    "nominal preferred over borrowing": dict(a=("nominal", 1, 0, None), b=("borrowing", 0, 1, None), want=True),
    "both borrowing at different levels falls through to FIFO": dict(a=("borrow-level-1", 1, 1, None), b=("borrow-level-2", 0, 2, None), want=False),
    "lower DRS preferred when both borrow": dict(a=("lower-drs", 1, 2, (-1.0, 1.0)), b=("higher-drs", 0, 1, (0.0, 0.0)), want=True),
    "both nominal falls through to FIFO": dict(a=("older", 0, 0, None), b=("newer", 1, 0, None), want=True),
}

# ---- TestSatisfiesPreemptionPolicy preemption/common/preemption_policy_test.go:32:
#      (preemptor priority, creation offset min), (candidate priority, creation offset min), policy, buffer gate, want
POLICY_CASES = {
    "LowerPriority: preemptor has higher priority": ((10, 0), (5, 0), "LowerPriority", False, True),
    "LowerPriority: preemptor has same priority": ((10, 0), (10, 0), "LowerPriority", False, False),
    "LowerOrNewerEqualPriority: preemptor has same priority, same timestamp": ((10, 0), (10, 0), "LowerOrNewerEqualPriority", False, False),
    "LowerOrNewerEqualPriority: preemptor has same priority, newer timestamp (within 5min buffer)": ((10, 1), (10, 0), "LowerOrNewerEqualPriority", False, False),
    "LowerOrNewerEqualPriority: preemptor has same priority, older timestamp (within 5min buffer)": ((10, 0), (10, 1), "LowerOrNewerEqualPriority", False, True),
    "LowerOrNewerEqualPriority with SchedulerTimestampPreemptionBuffer: ... older timestamp (within 5min buffer)": ((10, 0), (10, 1), "LowerOrNewerEqualPriority", True, False),
    "LowerOrNewerEqualPriority with SchedulerTimestampPreemptionBuffer: ... older timestamp (outside 5min buffer)": ((10, 0), (10, 6), "LowerOrNewerEqualPriority", True, True),
    "PreemptionPolicyAny": ((10, 0), (10, 6), "Any", False, True),
    "PreemptionPolicyNever": ((10, 0), (10, 6), "Never", False, False),
}

# ---- TestCandidatesOrdering preemption/preemption_test.go:4525 (the two AdmissionFairSharing cases are out of scope):
#      candidates (name, ClusterQueue, priority, quota reserved at offset s | None, evicted); preemptor CQ = "preemptor"
CANDIDATE_ORDER_CASES = {
    "workloads sorted by priority": dict(cands=[("high", "preemptor", 10, 0, False), ("low", "preemptor", -10, 0, False)], want=["low", "high"]),
    "evicted workload first": dict(cands=[("other", "preemptor", 10, 0, False), ("evicted", "other", 0, None, True)], want=["evicted", "other"]),
    "workload from different CQ first": dict(cands=[("preemptorCq", "preemptor", 10, 0, False), ("other", "other", 10, 0, False)], want=["other", "preemptorCq"]),
    "old workloads last": dict(cands=[("older", "preemptor", 0, -1, False), ("younger", "preemptor", 0, 1, False), ("current", "preemptor", 0, 0, False)],
                               want=["younger", "current", "older"]),
}

# ---- TestSearch flavorassigner/podset_reducer_test.go:26: podsets (count, minCount | None), countLimit -> (found, count)
#      ("empty" and "podset with replica count 0" have no pods to place and are not expressible as a workload)
REDUCER_CASES = {
    "partial not available": dict(podsets=[(1, None), (2, 2)], limit=2, found=False, count=0),
    "partial available": dict(podsets=[(5, 3), (5, 4), (5, 1), (5, 2)], limit=15, found=True, count=15),
    "one partial available": dict(podsets=[(5, 3), (5, None), (5, None), (5, None)], limit=19, found=True, count=19),
    "to min": dict(podsets=[(5, 3), (5, 4), (5, 1), (5, 2)], limit=10, found=True, count=10),
    "to max": dict(podsets=[(5, 3), (5, 4), (5, 1), (5, 2)], limit=20, found=True, count=20),
    "no overflow": dict(podsets=[(150_000, 1)] * 8, limit=150_000, found=True, count=150_000),
    "max pods on 1.27": dict(podsets=[(150_000, 1)] + [(1, None)] * 7, limit=150_000, found=True, count=150_000),
}

# ---- TestSnapshotAddRemoveWorkloadWithLendingLimit pkg/cache/scheduler/snapshot_test.go:1131
#      lend-a: cpu nominal 10, lendingLimit 4; lend-b: cpu nominal 10, lendingLimit 6 (cohort "lend");
#      workloads lend-a-1 (1 cpu), lend-a-2 (9), lend-a-3 (6) in lend-a, lend-b-1 (4) in lend-b.
#      remaining workloads after the case's remove/add -> Usage (milli-cpu) of cohort, lend-a, lend-b
LENDING_WORKLOADS = {"lend-a-1": ("lend-a", 1), "lend-a-2": ("lend-a", 9), "lend-a-3": ("lend-a", 6), "lend-b-1": ("lend-b", 4)}
LENDING_CASES = {
    "remove all": ([], (0, 0, 0)),
    "remove workload, but still using quota over GuaranteedQuota": (["lend-a-1", "lend-a-3", "lend-b-1"], (1_000, 7_000, 4_000)),
    "remove wokload, using same quota as GuaranteedQuota": (["lend-a-3", "lend-b-1"], (0, 6_000, 4_000)),
    "remove workload, using less quota than GuaranteedQuota": (["lend-a-1", "lend-b-1"], (0, 1_000, 4_000)),
    "remove all then add workload, using less quota than GuaranteedQuota": (["lend-a-1"], (0, 1_000, 0)),
    "remove all then add workload, using same quota as GuaranteedQuota": (["lend-a-3"], (0, 6_000, 0)),
    "remove all then add workload, using quota over GuaranteedQuota": (["lend-a-2"], (3_000, 9_000, 0)),
}

# ---- TestStrictFIFO pkg/cache/queue/cluster_queue_test.go:892: which of w1 / w2 the ClusterQueue pops first.
#      (creation offset s, priority, evicted-by-PodsReadyTimeout at offset s | None) x2, ordering, expected
HIGH, LOW = 1000, -1000
STRICT_FIFO_CASES = {
    "w1.priority is higher than w2.priority": ((0, HIGH, None), (1, LOW, None), "Eviction", "w1"),
    "w1.priority equals w2.priority and w1.create time is earlier than w2.create time": ((0, 0, None), (1, 0, None), "Eviction", "w1"),
    "... but w1 was evicted": ((0, 0, 2), (1, 0, None), "Eviction", "w2"),
    "... w1 was evicted but kueue is configured to always use the creation timestamp": ((0, 0, 2), (1, 0, None), "Creation", "w1"),
    "p1.priority is lower than p2.priority and w1.create time is earlier than w2.create time": ((0, LOW, None), (1, HIGH, None), "Eviction", "w2"),
}
