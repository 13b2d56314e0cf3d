"""Drain mode: iterate scheduling cycles over a snapshot whose pending tables hold whole queues.

This module is the HOST DEFINITION of the drain (numpy bookkeeping around any single-cycle evaluator); the product
path is kb_run_drain (include/kueue_b200.h), which keeps the queue layer on the device.  tests/test_drain.py runs this
loop on the oracle and compares kb_run_drain with it cycle by cycle.

  * per ClusterQueue the pending workloads are ordered once by queueOrderingFunc
    (pkg/cache/queue/cluster_queue.go:636-685: priority desc, queue-order timestamp, UID);
  * every cycle takes the current head of every ClusterQueue (queues.Heads, manager.go:770-794), runs ONE cycle and
    applies its decisions the way schedule() + requeueAndUpdate (scheduler.go:405-418,823-850) do:
      - Assumed: the workload leaves the queue, its Assignment.Usage is added to the ClusterQueue usage and it joins
        the admitted tables (a preemption candidate of later cycles);
      - every other entry keeps `LastAssignment = &assignment.LastState` (scheduler.go:494): the next attempt starts
        after the flavors already tried (ps_tried_idx -> ps_last_tried, ClusterQueueGeneration -> wl_last_gen);
        a Preempting entry gets LastAssignment = nil (scheduler.go:345);
      - StrictFIFO: the entry stays the head (requeue is immediate, cluster_queue.go:622-624);
      - BestEffortFIFO: skipped entries (FailedAfterNomination) and pending preemptions stay (immediate requeue);
        NoFit / Preempt-without-targets entries stay when LastAssignment.PendingFlavors() (cluster_queue.go:372,
        workload.go:163-176) and otherwise go to the inadmissible set: the next workload becomes the head;
      - BestEffortFIFO, mode NoFit, known SchedulingHash: every queued workload of the same equivalence class moves to
        the inadmissible set as well (handleInadmissibleHash, cluster_queue.go:408-425; scheduler.go:292-300);
  * the drain ends with the first cycle that admits nothing (no cluster event can change the outcome after that;
    evictions are asynchronous in the reference and are not replayed: a Preempting entry keeps its place and its
    targets stay admitted).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List

import numpy as np

from . import abi
from .shard import _csr_take


@dataclass
class DrainResult:
    cycles: int = 0
    admitted: List[np.ndarray] = field(default_factory=list)   # per cycle: global pending indices admitted
    heads: List[np.ndarray] = field(default_factory=list)      # per cycle: global pending indices evaluated
    decisions: List[np.ndarray] = field(default_factory=list)  # per cycle: KB_DEC_* per entry
    cq_usage: np.ndarray | None = None                         # final ClusterQueue usage [Q][FR]

    @property
    def n_admitted(self) -> int:
        return int(sum(len(a) for a in self.admitted))

    @property
    def n_decisions(self) -> int:
        return int(sum(len(h) for h in self.heads))


# RequeueReason (pkg/cache/queue/cluster_queue.go:55-66) of an entry that was not admitted, from the cycle's decision:
# entries that were nominated and then lost the admit loop are "FailedAfterNomination" (requeueAndUpdate
# scheduler.go:825-828), a Preempting entry waits for its evictions ("PendingPreemption", scheduler.go:339-346), NoFit and
# Preempt-without-targets entries keep the generic reason.
REASON_GENERIC, REASON_FAILED_AFTER_NOMINATION, REASON_NAMESPACE_MISMATCH, REASON_PENDING_PREEMPTION = (
    "", "FailedAfterNomination", "NamespaceMismatch", "PendingPreemption")


def requeue_reason(decision: int) -> str:
    if decision in (abi.DEC_NOFIT, abi.DEC_PREEMPT_NO_TARGETS):
        return REASON_GENERIC
    if decision == abi.DEC_PREEMPTING:
        return REASON_PENDING_PREEMPTION
    return REASON_FAILED_AFTER_NOMINATION  # skipped in the admit loop


def pending_flavors(last_tried) -> bool:
    """AssignmentClusterQueueState.PendingFlavors (workload.go:163-176): some podset resource still has a flavor to
    try.  `last_tried`: LastTriedFlavorIdx as rows of per-resource indices (-1 = exhausted / absent), or None."""
    if last_tried is None:
        return False
    return bool((np.asarray(last_tried) != -1).any())


def requeue_goes_inadmissible(strict_fifo: bool, reason: str, last_tried) -> bool:
    """RequeueIfNotPresent + requeueIfNotPresent (cluster_queue.go:362-393,609-633) inside a drain: no
    QueueInadmissibleWorkloads call happens between Pop and the requeue and no backoff is pending, so the workload
    returns to the heap iff the requeue is immediate or flavors are pending; otherwise it joins the inadmissible set."""
    if strict_fifo:
        immediate = reason != REASON_NAMESPACE_MISMATCH
    else:
        immediate = reason in (REASON_FAILED_AFTER_NOMINATION, REASON_PENDING_PREEMPTION, "PreemptionFailed")
    return not (immediate or pending_flavors(last_tried))


def _queue_order(a, Q):
    """Pending workloads of every ClusterQueue in queueOrderingFunc order: (order, start[Q+1])."""
    cq = a["wl_cq"].astype(np.int64)
    order = np.lexsort((a["wl_uid"], a["wl_ts"], -a["wl_priority"].astype(np.int64), cq))
    start = np.zeros(Q + 1, np.int64)
    np.add.at(start, cq + 1, 1)
    return order, np.cumsum(start)


def drain(snap: abi.FlatSnapshot, run_cycle: Callable[[abi.FlatSnapshot], abi.CycleOut], max_cycles: int = 10_000) -> DrainResult:
    a = snap.arrays
    Q, R, FR = snap.n_cq, snap.n_resource, snap.n_fr
    order, qstart = _queue_order(a, Q)
    qlen = np.diff(qstart)
    cursor = np.zeros(Q, np.int64)
    strict = a["cq_strategy"].astype(bool)  # KB_QUEUE_STRICT_FIFO = 1
    usage = a["cq_usage"].reshape(Q, FR).astype(np.int64).copy()
    adm = {k: a[k].copy() for k in ("adm_cq", "adm_priority", "adm_ts", "adm_qr_ts", "adm_uid", "adm_evicted", "adm_use_start",
                                    "adm_use_fr", "adm_use_qty")}
    ps_start = a["wl_ps_start"].astype(np.int64)
    ps_req = a["ps_req"].reshape(-1, R)
    res = DrainResult()
    last_gen = a["wl_last_gen"].astype(np.int64).copy()          # LastAssignment.ClusterQueueGeneration, -1 = nil
    last_tried = a["ps_last_tried"].reshape(-1, R).copy()         # LastAssignment.LastTriedFlavorIdx
    shash = a.get("wl_sched_hash")
    has_qr = a.get("wl_has_quota_reservation")
    gone = np.zeros(snap.n_wl, bool)                              # moved to the inadmissible set by its scheduling hash
    static = {k: v for k, v in a.items() if not (k.startswith(("wl_", "ps_", "adm_")) or k in ("heads", "cq_usage"))}
    covers_pods = np.zeros(Q, bool)
    if snap.pods_resource >= 0:
        for q in range(Q):
            masks = a["rg_res_mask"][a["cq_rg_start"][q]:a["cq_rg_start"][q + 1]]
            covers_pods[q] = bool(np.any(masks & (1 << snap.pods_resource)))
    for cyc in range(max_cycles):
        live = np.flatnonzero(cursor < qlen)
        if len(live) == 0:
            break
        heads = order[qstart[live] + cursor[live]]  # global pending indices, one per ClusterQueue with work left
        cs = abi.FlatSnapshot(n_cq=Q, n_cohort=snap.n_cohort, n_flavor=snap.n_flavor, n_resource=R, pods_resource=snap.pods_resource,
                              flags=snap.flags, now_ns=snap.now_ns + cyc, static_generation=snap.static_generation)
        cs.arrays.update(static)
        cs.set("cq_usage", usage)
        for k, v in adm.items():
            cs.set(k, v)
        for nm in ("wl_cq", "wl_priority", "wl_ts", "wl_uid"):
            cs.set(nm, a[nm][heads])
        cs.set("wl_last_gen", last_gen[heads])
        if has_qr is not None:
            cs.set("wl_has_quota_reservation", has_qr[heads])
        st, rows = _csr_take(ps_start, heads)
        cs.set("wl_ps_start", st)
        cs.set("ps_req", ps_req[rows]); cs.set("ps_last_tried", last_tried[rows])
        for nm in ("ps_req_mask", "ps_count", "ps_min_count", "ps_flavor_ok"):
            cs.set(nm, a[nm][rows])
        if "ps_group" in a:  # optional tables (kb_snapshot: NULL when absent)
            cs.set("ps_group", a["ps_group"][rows])
        cs.set("heads", np.arange(len(heads)))
        cs.finalize()
        out = run_cycle(cs)
        dec = np.asarray(out.decision).copy()
        res.cycles += 1
        res.heads.append(heads); res.decisions.append(dec)
        ok = dec == abi.DEC_ASSUMED
        res.admitted.append(heads[ok])
        # ---- apply the admissions: Assignment.Usage (flavorassigner.go:198-218) with the admitted counts
        new_fr, new_qty, new_start = [], [], [int(adm["adm_use_start"][-1])]
        for e in np.flatnonzero(ok):
            cqi = int(cs.arrays["wl_cq"][e])
            cells = {}
            for row in range(int(st[e]), int(st[e + 1])):
                full, cnt = int(cs.arrays["ps_count"][row]), int(out.ps_count[row])
                for r in range(R):
                    f = int(out.ps_flavor[row, r])
                    if f < 0:
                        continue
                    if covers_pods[cqi] and r == snap.pods_resource:
                        q = cnt
                    else:
                        q = int(cs.arrays["ps_req"].reshape(-1, R)[row, r])
                        if full != 0 and full != cnt:
                            q = q // full * cnt  # ScaledTo workload.go:258-275
                    cells[f * R + r] = cells.get(f * R + r, 0) + q
            for fr, q in cells.items():
                usage[cqi, fr] += q
                new_fr.append(fr); new_qty.append(q)
            new_start.append(new_start[-1] + len(cells))
        n_new = int(ok.sum())
        if n_new:
            g = heads[ok]
            adm["adm_cq"] = np.concatenate([adm["adm_cq"], a["wl_cq"][g]])
            adm["adm_priority"] = np.concatenate([adm["adm_priority"], a["wl_priority"][g]])
            adm["adm_ts"] = np.concatenate([adm["adm_ts"], a["wl_ts"][g]])
            adm["adm_qr_ts"] = np.concatenate([adm["adm_qr_ts"], np.full(n_new, snap.now_ns + cyc, np.int64)])
            adm["adm_uid"] = np.concatenate([adm["adm_uid"], a["wl_uid"][g]])
            adm["adm_evicted"] = np.concatenate([adm["adm_evicted"], np.zeros(n_new, np.uint8)])
            adm["adm_use_start"] = np.concatenate([adm["adm_use_start"], np.asarray(new_start[1:], np.int32)])
            adm["adm_use_fr"] = np.concatenate([adm["adm_use_fr"], np.asarray(new_fr, np.int32)])
            adm["adm_use_qty"] = np.concatenate([adm["adm_use_qty"], np.asarray(new_qty, np.int64)])
        # ---- LastAssignment of the entries that stay pending (scheduler.go:345,494)
        tried = np.asarray(out.ps_tried_idx).reshape(-1, R)
        cqs = live
        inadmissible = np.zeros(len(heads), bool)
        for e in range(len(heads)):
            if ok[e]:
                continue
            w = int(heads[e])
            r0, r1 = int(st[e]), int(st[e + 1])
            if dec[e] == abi.DEC_PREEMPTING:
                last_gen[w] = -1
                last = None
            else:
                last_gen[w] = int(a["cq_generation"][int(a["wl_cq"][w])])
                last_tried[int(ps_start[w]):int(ps_start[w + 1])] = tried[r0:r1]
                last = tried[r0:r1]
            # ---- queue movement (cluster_queue.go:362-425,609-633); a StrictFIFO queue keeps its head either way
            inadmissible[e] = requeue_goes_inadmissible(False, requeue_reason(int(dec[e])), last)
        advance = ok | (inadmissible & ~strict[cqs])
        if shash is not None:
            for e in np.flatnonzero((dec == abi.DEC_NOFIT) & ~strict[cqs]):
                hsh = int(shash[heads[e]])
                if hsh == 0:
                    continue
                q = int(cqs[e])
                rest = order[qstart[q] + cursor[q] + 1:qstart[q + 1]]
                gone[rest[shash[rest] == hsh]] = True
        cursor[cqs[advance]] += 1
        for q in cqs:  # the next head is the first queued workload that was not set aside
            while cursor[q] < qlen[q] and gone[order[qstart[q] + cursor[q]]]:
                cursor[q] += 1
        if n_new == 0:
            break
    res.cq_usage = usage
    return res
