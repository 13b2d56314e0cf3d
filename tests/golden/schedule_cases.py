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
