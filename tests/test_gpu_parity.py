"""GPU parity: the CUDA path through the C-ABI vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

import oracle
from kueue_b200 import abi, synth
from kueue_b200.api import flatten
from tests.golden.tree_cases import TREE_CASES
from tests.helpers import assert_cycle_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ev():
    from kueue_b200 import native
    e = native.Evaluator(0)
    yield e
    e.close()


@pytest.mark.parametrize("name", list(TREE_CASES))
def test_tree_golden(ev, name):
    tc = TREE_CASES[name]
    snap, idx = flatten(tc["cqs"], tc["cohorts"], usage=tc["usage"])
    out = ev.tree_eval(snap)
    for cq, m in tc["want_available"].items():
        for (f, r), v in m.items():
            assert out.available[idx.cqs.index(cq), idx.fr(f, r)] == v
    for cq, m in tc["want_potential"].items():
        for (f, r), v in m.items():
            assert out.potential_available[idx.cqs.index(cq), idx.fr(f, r)] == v


@pytest.mark.parametrize("config,kw", [(1, {}), (2, dict(W=5000, Q=50)), (2, dict(W=20000, Q=200, podsets_max=3)),
                                        (4, dict(W=4000, Q=200))])
def test_tree_matches_oracle(ev, config, kw):
    snap = synth.make_snapshot(config, **kw)
    got, want = ev.tree_eval(snap), oracle.tree_eval(snap)
    for f in ("subtree_quota", "usage", "available", "potential_available", "drs_rounded", "drs_resource", "drs_borrowing"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f


@pytest.mark.parametrize("config,kw", [
    (1, {}), (1, dict(heads="one_per_cq")),
    (2, dict(W=5000, Q=50)), (2, dict(W=5000, Q=50, heads="one_per_cq")),
    (2, dict(W=20000, Q=200, podsets_max=3)),
    (4, dict(W=4000, Q=200)), (4, dict(W=4000, Q=200, heads="one_per_cq", podsets_max=2)),
])
def test_cycle_matches_oracle(ev, config, kw):
    snap = synth.make_snapshot(config, **kw)
    got, want = ev.run_cycle(snap), oracle.run_cycle(snap)
    assert_cycle_equal(got, want)


def _fair(snap):
    snap.flags |= abi.F_FAIR_SHARING
    return snap


@pytest.mark.parametrize("make", [
    lambda: synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq"),
    lambda: synth.make_snapshot(3, W=30000, Q=1000, heads="one_per_cq", podsets_max=2),
    lambda: _fair(synth.make_snapshot(4, W=4000, Q=200, heads="one_per_cq", preemption=False)),      # fair sharing over a depth-4 tree
    lambda: _fair(synth.make_snapshot(4, W=40000, Q=2000, heads="one_per_cq", preemption=False)),    # tree too large for shared memory
    lambda: _fair(synth.make_snapshot(2, W=2000, Q=100, heads="one_per_cq")),      # fair sharing, CQs without cohorts
])
def test_fair_sharing_cycle_matches_oracle(ev, make):
    snap = make()
    got, want = ev.run_cycle(snap), oracle.run_cycle(snap)
    assert_cycle_equal(got, want)


def test_full_size_config3_single_cycle(ev):
    snap = synth.make_snapshot(3, heads="one_per_cq")
    got, want = ev.run_cycle(snap), oracle.run_cycle(snap)
    assert_cycle_equal(got, want)


def _classical(snap):
    snap.flags &= ~abi.F_FAIR_SHARING
    return snap


@pytest.mark.parametrize("make", [
    lambda: synth.make_snapshot(2, W=5000, Q=50, preemption=True, tight=1.2),                      # within-ClusterQueue preemption
    lambda: synth.make_snapshot(2, W=5000, Q=50, preemption=True, tight=1.2, podsets_max=3, seed=9),
    lambda: _classical(synth.make_snapshot(3, W=3000, Q=300, preemption=True, heads="one_per_cq", tight=1.1)),  # flat cohorts: reclaim
    lambda: synth.make_snapshot(4, W=4000, Q=200, heads="one_per_cq", tight=1.1),                  # depth-4 tree, heavy preemption
    lambda: synth.make_snapshot(4, W=2000, Q=100, tight=1.06),                                      # batched heads
    lambda: synth.make_snapshot(4, W=40000, Q=2000, heads="one_per_cq"),                            # trees too large for shared memory
])
def test_preemption_cycle_matches_oracle(ev, make):
    snap = make()
    cap = 40 * snap.n_adm + 10000
    got, want = ev.run_cycle(snap, abi.CycleOut(snap, cap)), oracle.run_cycle(snap, cap)
    assert want.n_targets > 0
    assert_cycle_equal(got, want)


@pytest.mark.parametrize("make", [
    lambda: synth.make_snapshot(2, W=5000, Q=50, partial=True, podsets_max=2),
    lambda: synth.make_snapshot(4, W=2000, Q=100, partial=True, podsets_max=3, tight=1.06),
])
def test_partial_admission_matches_oracle(ev, make):
    snap = make()
    cap = 40 * snap.n_adm + 10000
    got, want = ev.run_cycle(snap, abi.CycleOut(snap, cap)), oracle.run_cycle(snap, cap)
    assert (want.ps_count != snap.ps_count).any(), "fixture must exercise reduced counts"
    assert_cycle_equal(got, want)


def _grouped(snap, seed=3):
    """Random PodSetGroups: in every workload with >= 2 podsets the first two (sometimes three) rows share a group."""
    rng = np.random.default_rng(seed)
    st = np.asarray(snap.arrays["wl_ps_start"])
    grp = np.full(snap.n_podset, -1, np.int32)
    for w in np.flatnonzero(np.diff(st) >= 2):
        if rng.random() < 0.7:
            n = min(int(st[w + 1] - st[w]), 2 + int(rng.random() < 0.3))
            grp[st[w]:st[w] + n] = 0
    assert (grp >= 0).any()
    snap.set("ps_group", grp)
    return snap.finalize() if hasattr(snap, "finalize") else snap


@pytest.mark.parametrize("make", [
    lambda: synth.make_snapshot(2, W=20000, Q=200, podsets_max=3),                                             # k_nominate
    lambda: synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq", podsets_max=3),                          # fused per-root cycle
    lambda: synth.make_snapshot(2, W=5000, Q=50, preemption=True, tight=1.2, podsets_max=3, seed=9),           # walk with searches
    lambda: synth.make_snapshot(4, W=4000, Q=200, heads="one_per_cq", tight=1.1, podsets_max=3),
    lambda: synth.make_snapshot(4, W=2000, Q=100, partial=True, podsets_max=3, tight=1.06),                    # reducer x groups
    lambda: _fair(synth.make_snapshot(3, W=3000, Q=300, preemption=True, heads="one_per_cq", tight=1.1, podsets_max=3)),
])
def test_podset_groups_match_oracle(ev, make):
    """PodSetGroup units (flavorassigner.go:613-675): grouped podsets get one flavor search over their summed requests."""
    snap = _grouped(make())
    cap = 40 * snap.n_adm + 10000
    got, want = ev.run_cycle(snap, abi.CycleOut(snap, cap)), oracle.run_cycle(snap, cap)
    plain = oracle.run_cycle(_ungrouped(snap), cap)
    assert not np.array_equal(plain.ps_flavor, want.ps_flavor) or not np.array_equal(plain.decision, want.decision), "groups must matter in the fixture"
    assert_cycle_equal(got, want)


def _ungrouped(snap):
    import copy
    s2 = copy.copy(snap); s2.arrays = dict(snap.arrays); s2.arrays.pop("ps_group", None); s2._struct = None
    return s2


def _fair_golden_cases():
    import json, os
    from tests.test_oracle_golden_preemption2 import fair_flags
    here = os.path.dirname(__file__)
    out = []
    for name, tc in json.load(open(os.path.join(here, "golden", "preemption_fair_cases.json"))).items():
        out.append(pytest.param((tc, fair_flags(tc)), id=f"fair:{name[:60]}"))
    return out


@pytest.mark.parametrize("tcf", _fair_golden_cases())
def test_reference_fair_preemption_scenarios_cycle(ev, tcf):
    """TestFairPreemptions scenarios (preemption_fair_test.go:46) as one fair-sharing cycle: device vs oracle."""
    from tests.golden_loader import build_preemption_case
    tc, flags = tcf
    snap, idx = build_preemption_case(tc, flags)
    got, want = ev.run_cycle(snap), oracle.run_cycle(snap)
    assert_cycle_equal(got, want)


@pytest.mark.parametrize("make", [
    lambda: synth.make_snapshot(3, W=3000, Q=300, preemption=True, heads="one_per_cq", tight=1.1),
    lambda: synth.make_snapshot(3, W=3000, Q=300, preemption=True, heads="one_per_cq", tight=1.05, seed=11, podsets_max=2),
    lambda: _fair(synth.make_snapshot(4, W=4000, Q=200, heads="one_per_cq", tight=1.1)),
    lambda: _fair(synth.make_snapshot(4, W=40000, Q=2000, heads="one_per_cq")),
])
def test_fair_preemption_cycle_matches_oracle(ev, make):
    snap = make()
    cap = 40 * snap.n_adm + 10000
    got, want = ev.run_cycle(snap, abi.CycleOut(snap, cap)), oracle.run_cycle(snap, cap)
    assert want.n_targets > 0
    assert_cycle_equal(got, want)


def _assign_golden_cases():
    import json, os
    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "assign_flavors_cases.json")))
    return [pytest.param(tc, id=f"assign:{name[:60]}") for name, tc in d["cases"].items()]


@pytest.mark.parametrize("tc", _assign_golden_cases())
def test_reference_assign_flavors_scenarios_cycle(ev, tc):
    """TestAssignFlavors snapshots (flavorassigner_test.go:165) as one cycle: device vs oracle (both with the real
    preemption search instead of the test's stub)."""
    from tests.test_oracle_golden_assign import build
    snap, idx = build(tc)
    got, want = ev.run_cycle(snap), oracle.run_cycle(snap)
    assert_cycle_equal(got, want)


def _golden_cases():
    import json, os
    here = os.path.dirname(__file__)
    out = []
    for f in ("preemption_cases.json", "preemption_hierarchical_cases.json"):
        for name, tc in json.load(open(os.path.join(here, "golden", f))).items():
            out.append(pytest.param(tc, id=f"{f.split('_cases')[0]}:{name[:60]}"))
    return out


@pytest.mark.parametrize("tc", _golden_cases())
def test_reference_preemption_scenarios_cycle(ev, tc):
    """The reference's TestPreemption / TestHierarchicalPreemptions scenarios as one scheduling cycle
    with the incoming workload as the only head: device vs oracle, bit-exact."""
    from tests.golden_loader import build_preemption_case
    snap, idx = build_preemption_case(tc)
    got, want = ev.run_cycle(snap), oracle.run_cycle(snap)
    assert_cycle_equal(got, want)


def test_static_generation_reuse(ev):
    """kb_snapshot.static_generation: unchanged generation => quota/topology tables are not re-read; a bumped
    generation picks up new quotas."""
    snap = synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq")
    snap.static_generation = 41
    want = oracle.run_cycle(snap)
    assert_cycle_equal(ev.run_cycle(snap), want)
    h2d_full = ev.stats().h2d_bytes
    # only the usage changes between cycles
    snap.arrays["cq_usage"][:] = (snap.arrays["cq_usage"] * 0.9).astype(np.int64)
    want2 = oracle.run_cycle(snap)
    assert_cycle_equal(ev.run_cycle(snap), want2)
    assert ev.stats().h2d_bytes < h2d_full - 3 * snap.n_nodes * snap.n_fr * 8 + 1, "static tables were uploaded again"
    # a spec change comes with a new generation
    snap.arrays["nominal"][:] = (snap.arrays["nominal"] * 1.5).astype(np.int64)
    snap.static_generation = 42
    assert_cycle_equal(ev.run_cycle(snap), oracle.run_cycle(snap))


def test_full_size_config2(ev):
    snap = synth.make_snapshot(2)
    got, want = ev.run_cycle(snap), oracle.run_cycle(snap)
    assert_cycle_equal(got, want)


def _entry_ordering_cases():
    from tests.golden.schedule_cases import ENTRY_ORDERING_CASES
    return [pytest.param(tc, id=f"order:{name[:60]}") for name, tc in ENTRY_ORDERING_CASES.items()]


@pytest.mark.gpu
@pytest.mark.parametrize("tc", _entry_ordering_cases())
def test_reference_entry_ordering_cycle(ev, tc):
    """TestEntryOrdering (scheduler_test.go:8291) replayed as one cycle: the device's commit order is the
    reference's wantOrder (and equals the oracle's)."""
    from tests.schedule_golden import entry_ordering_snapshot, order_of
    snap, idx, want_borrow = entry_ordering_snapshot(tc)
    got, want = ev.run_cycle(snap), oracle.run_cycle(snap)
    assert_cycle_equal(got, want)
    assert (got.borrow == want_borrow).all()
    assert order_of(got, tc) == tc["want"]


def _schedule_cycle_cases():
    from tests.test_oracle_golden_schedule_cycle import schedule_cases
    return [pytest.param(n, id=f"schedule:{n[:60]}") for n in schedule_cases()]


@pytest.mark.parametrize("name", _schedule_cycle_cases())
def test_reference_schedule_scenarios_cycle(ev, name):
    """TestSchedule (scheduler_test.go:69) replayed on the device: the reference's wantAssignments / preempted
    workloads / skipped preemptions, and bit-exact agreement with the oracle."""
    from tests.test_oracle_golden_schedule_cycle import DOC, check_schedule_case
    snap, got = check_schedule_case(DOC["cases"][name], ev.run_cycle)
    assert_cycle_equal(got, oracle.run_cycle(snap))


def test_span_upload_from_one_pinned_block(ev):
    """Tables carved out of one kb_alloc_pinned block travel in a single DMA (span upload) — same decisions."""
    from kueue_b200 import native
    for make in (lambda: synth.make_snapshot(3, W=3000, Q=300, preemption=True, heads="one_per_cq", tight=1.1),
                 lambda: synth.make_snapshot(4, W=4000, Q=200, heads="one_per_cq", tight=1.1),
                 lambda: synth.make_snapshot(2, W=5000, Q=50, podsets_max=3)):
        snap = make()
        pinned = native.pin_snapshot(snap)
        cap = 40 * snap.n_adm + 10000
        want = oracle.run_cycle(snap, cap)
        got = ev.run_cycle(pinned, abi.CycleOut(pinned, cap))
        assert_cycle_equal(got, want)
        pinned.static_generation = 7
        assert_cycle_equal(ev.run_cycle(pinned, abi.CycleOut(pinned, cap)), want)
        assert_cycle_equal(ev.run_cycle(pinned, abi.CycleOut(pinned, cap)), want)  # static tables reused, span upload again
        # results into one kb_alloc_cycle_out block: the per-entry / per-podset tables come back with a single DMA
        pout = native.pin_cycle_out(abi.CycleOut(pinned, cap))
        assert_cycle_equal(ev.run_cycle(pinned, pout), want)
        assert ev.stats().d2h_bytes > 0
        assert_cycle_equal(ev.run_cycle(pinned, pout), want)


def test_reference_podset_reducer_cycle(ev):
    """TestSearch (podset_reducer_test.go:26) inside the cycle: device == oracle, total admitted pods == wantCount."""
    from tests.golden.schedule_cases import REDUCER_CASES
    from tests.schedule_golden import reducer_snapshot
    for name, tc in REDUCER_CASES.items():
        snap, idx = reducer_snapshot(tc)
        got, want = ev.run_cycle(snap), oracle.run_cycle(snap)
        assert_cycle_equal(got, want)
        if tc["found"]:
            assert got.decision[0] == abi.DEC_ASSUMED and int(got.ps_count.sum()) == tc["count"], name


def test_reference_assign_in_cohorts_cycle(ev):
    """TestHierarchical / TestReclaimBeforePriorityPreemption environments (no stub oracle here): device == oracle."""
    from tests.test_oracle_golden_assign_extra import DATA, build_assign_extra
    for func, tab in DATA.items():
        for name, tc in tab.items():
            snap, idx = build_assign_extra(func, tc)
            assert_cycle_equal(ev.run_cycle(snap), oracle.run_cycle(snap))


@pytest.mark.parametrize("make", [
    lambda: synth.make_snapshot(2, W=3000, Q=40, F=20, R=8, podsets_max=2),                          # FR = 160: no register / two-column fast paths
    lambda: synth.make_snapshot(3, W=600, Q=60, F=12, R=8, heads="one_per_cq", preemption=True, tight=1.1),  # FR = 96, fair
    lambda: synth.make_snapshot(4, W=2000, Q=100, F=20, R=8, heads="one_per_cq", tight=1.25),         # FR = 160 > KB_MAX_CELLS: no speculative searches
    lambda: synth.make_snapshot(4, W=2000, Q=100, F=10, R=8, heads="one_per_cq", tight=1.25),         # FR = 80: speculative, three cells per lane
])
def test_wide_flavor_resource_tables(ev, make):
    snap = make()
    cap = 40 * snap.n_adm + 10000
    assert_cycle_equal(ev.run_cycle(snap, abi.CycleOut(snap, cap)), oracle.run_cycle(snap, cap))


def test_reference_usage_with_lending_limit(ev):
    """TestSnapshotAddRemoveWorkloadWithLendingLimit (snapshot_test.go:1131): k_tree's cohort usage."""
    from tests.golden.schedule_cases import LENDING_CASES
    from tests.schedule_golden import lending_snapshot
    for name, (remaining, (cohort, a, b)) in LENDING_CASES.items():
        snap, idx = lending_snapshot(remaining)
        out = ev.tree_eval(snap)
        fr = idx.fr("default", "cpu")
        assert (int(out.usage[idx.node("lend"), fr]), int(out.usage[idx.node("lend-a"), fr]), int(out.usage[idx.node("lend-b"), fr])) == (cohort, a, b), name


def test_reference_last_assignment_outdated_cycle(ev):
    """TestLastAssignmentOutdated (flavorassigner_test.go:3705) through the cycle: device == oracle."""
    from tests.test_oracle_golden_schedule import _last_assignment_snapshot
    for cq_gen, wl_gen, flavor in ((1, 0, "one"), (0, 0, "two")):
        snap, idx = _last_assignment_snapshot(cq_gen, wl_gen)
        got = ev.run_cycle(snap)
        assert_cycle_equal(got, oracle.run_cycle(snap))
        assert idx.flavors[int(got.ps_flavor[0, idx.resources.index("cpu")])] == flavor


@pytest.mark.parametrize("make", [
    lambda: synth.make_snapshot(3, W=900, Q=90, F=12, R=4, heads="one_per_cq"),                 # FR = 48: two columns per lane in the flat commit loop
    lambda: _classical(synth.make_snapshot(3, W=900, Q=90, F=12, R=4, heads="one_per_cq")),
    lambda: _classical(synth.make_snapshot(3, W=3000, Q=60, F=16, R=4)),                         # FR = 64, several entries per ClusterQueue (stale prefetch path)
    lambda: _classical(synth.make_snapshot(3, W=3000, Q=60)),                                    # FR = 32, batched heads: same-CQ neighbours
])
def test_flat_commit_loop_variants(ev, make):
    snap = make()
    assert_cycle_equal(ev.run_cycle(snap), oracle.run_cycle(snap))


@pytest.mark.parametrize("fair", [False, True])
def test_mixed_sign_priorities_in_one_cohort(ev, fair):
    """Priorities of both signs (incl. INT32_MIN / INT32_MAX) inside one root cohort: the iterator key must order
    them as signed ints (scheduler.go:799-805, fair_sharing_iterator.go:184-190) on the classical and the
    fair flat-cohort paths."""
    snap = synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq")
    if not fair:
        snap = _classical(snap)
    rng = np.random.default_rng(5)
    snap.arrays["wl_priority"][:] = rng.choice(np.array([-5, 0, 7, -2**31, 2**31 - 1, -1, 1], np.int64), snap.n_wl).astype(np.int32)
    # identical timestamps inside a cohort would hide a priority mix-up behind the timestamp key
    got, want = ev.run_cycle(snap), oracle.run_cycle(snap)
    assert_cycle_equal(got, want)
    # and with lone ClusterQueues holding several entries each
    snap = synth.make_snapshot(2, W=5000, Q=50)
    snap.arrays["wl_priority"][:] = rng.choice(np.array([-5, 0, 7, -2**31, 2**31 - 1], np.int64), snap.n_wl).astype(np.int32)
    assert_cycle_equal(ev.run_cycle(snap), oracle.run_cycle(snap))


@pytest.fixture(scope="module")
def cfg4_full():
    """BASELINE cfg4 at full size (Q=10 000, C=1 310, A=200 000, 10 roots x 1 131 nodes) and the oracle's cycle."""
    snap = synth.make_snapshot(4, heads="one_per_cq")
    cap = 40 * snap.n_adm
    return snap, cap, oracle.run_cycle(snap, cap)


def test_full_size_config4_single_cycle(ev, cfg4_full):
    snap, cap, want = cfg4_full
    assert want.n_targets > 0
    got = ev.run_cycle(snap, abi.CycleOut(snap, cap))
    assert_cycle_equal(got, want)


@pytest.mark.parametrize("world", [2, 8])
def test_full_size_config4_sharded_device(ev, cfg4_full, world):
    """shard.shard -> device cycle per shard -> shard.merge == the unsharded oracle (the N>1 path of bench.py with
    the CUDA evaluator in the loop)."""
    from kueue_b200 import shard
    snap, cap, want = cfg4_full
    node_rank = shard.partition_roots(snap, world)
    parts = []
    for r in range(world):
        sub, m = shard.shard(snap, r, world, node_rank)
        parts.append((ev.run_cycle(sub, abi.CycleOut(sub, cap)), m))
    assert_cycle_equal(shard.merge(snap, parts), want)


@pytest.mark.parametrize("world", [2, 8])
def test_config3_sharded_device(ev, world):
    from kueue_b200 import shard
    snap = synth.make_snapshot(3, W=200_000, Q=2000, heads="one_per_cq")
    want = oracle.run_cycle(snap)
    node_rank = shard.partition_roots(snap, world)
    parts = []
    for r in range(world):
        sub, m = shard.shard(snap, r, world, node_rank)
        parts.append((ev.run_cycle(sub), m))
    assert_cycle_equal(shard.merge(snap, parts), want)


def _last_context_cases():
    from tests.last_context_golden import DOC
    return [pytest.param(n, id=f"lastctx:{n[:60]}") for n in DOC["cases"]]


@pytest.mark.parametrize("name", _last_context_cases())
def test_reference_last_scheduling_context_two_cycles(ev, name):
    """TestLastSchedulingContext (scheduler_test.go:8569) on the device: two cycles with ps_tried_idx fed back as
    ps_last_tried; the reference's admissions after the second cycle, and device == oracle in both cycles."""
    from tests.last_context_golden import DOC, check
    outs = check(DOC["cases"][name], ev.run_cycle)
    for snap, got in outs:
        assert_cycle_equal(got, oracle.run_cycle(snap))


def _with_qr(snap, seed=11):
    rng = np.random.default_rng(seed)
    snap.set("wl_has_quota_reservation", (rng.random(snap.n_wl) < 0.3).astype(np.uint8))
    return snap.finalize() if hasattr(snap, "finalize") else snap


@pytest.mark.parametrize("v1", [False, True], ids=["k_cycle_flat", "k_cycle_root"])
@pytest.mark.parametrize("make", [
    lambda: synth.make_snapshot(3, W=900, Q=90, F=3, R=3, heads="one_per_cq"),                             # FR = 9: odd rows, columns not aligned to the CTA
    lambda: synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq", podsets_max=3, partial=True),        # PodSetReducer on the relocated tables
    lambda: _grouped(synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq", podsets_max=3)),            # PodSetGroup units
    lambda: synth.make_snapshot(3, W=1200, Q=120, F=16, R=4, heads="one_per_cq"),                          # FR = 64
    lambda: _classical(_with_qr(synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq"))),               # HasQuotaReservation key
    lambda: _classical(synth.make_snapshot(3, W=2000, Q=200, F=2, R=2, heads="one_per_cq")),               # FR = 4: rows shorter than a warp
    lambda: synth.make_snapshot(3, W=30, Q=3, heads="one_per_cq"),                                         # one tiny root
])
def test_fused_flat_kernels(ev, monkeypatch, make, v1):
    """Both fused per-root kernels on flat cohorts: k_cycle_flat (kb_flat.cuh, the relocated per-root snapshot in shared
    memory) and k_cycle_root (KB_FUSED_V1: the kernel it falls back to when the relocated copy does not fit)."""
    if v1:
        monkeypatch.setenv("KB_FUSED_V1", "1")
    snap = make()
    assert_cycle_equal(ev.run_cycle(snap), oracle.run_cycle(snap))


def test_incremental_usage_rows(ev):
    """kb_snapshot.usage_delta_* (SURVEY f2): after one call that leaves the usage table resident (KB_F_USAGE_RESIDENT),
    cycles that pass only the changed ClusterQueue rows decide exactly like cycles that pass the full table."""
    from kueue_b200 import native
    rng = np.random.default_rng(17)
    base = synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq")
    base.static_generation = 41
    base.flags |= abi.F_USAGE_RESIDENT
    assert_cycle_equal(ev.run_cycle(base), oracle.run_cycle(base))
    usage = np.array(base.arrays["cq_usage"]).reshape(base.n_cq, base.n_fr).copy()
    for step in range(3):
        dirty = np.sort(rng.choice(base.n_cq, size=[40, 1, 300][step], replace=False)).astype(np.int32)
        usage[dirty] = (usage[dirty] * rng.uniform(0.3, 1.4, (len(dirty), base.n_fr))).astype(np.int64)
        full = synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq")   # what the oracle sees: the whole patched table
        full.set("cq_usage", usage)
        inc = synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq")    # what the library gets: only the rows
        inc.static_generation = 41
        inc.set("usage_delta_cq", dirty); inc.set("usage_delta_rows", usage[dirty])
        inc.arrays["cq_usage"] = np.zeros(0, np.int64)                     # not read
        got = ev.run_cycle(inc, abi.CycleOut(full))
        assert_cycle_equal(got, oracle.run_cycle(full))
    # an empty delta list re-evaluates the resident table
    inc = synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq"); inc.static_generation = 41
    inc.set("usage_delta_cq", np.zeros(0, np.int32)); inc.set("usage_delta_rows", np.zeros(0, np.int64))
    inc.arrays["cq_usage"] = np.zeros(0, np.int64)
    assert_cycle_equal(ev.run_cycle(inc, abi.CycleOut(full)), oracle.run_cycle(full))
    # errors: a duplicate row, a stale generation, no resident table
    bad = synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq"); bad.static_generation = 41
    bad.set("usage_delta_cq", np.array([5, 5], np.int32)); bad.set("usage_delta_rows", usage[[5, 5]])
    with pytest.raises(native.KueueB200Error):
        ev.run_cycle(bad)
    bad.static_generation = 42
    bad.set("usage_delta_cq", np.array([5], np.int32)); bad.set("usage_delta_rows", usage[[5]])
    with pytest.raises(native.KueueB200Error):
        ev.run_cycle(bad)
    plain = synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq"); plain.static_generation = 41
    ev.run_cycle(plain)                                 # a full table without the hint drops the resident one
    bad.static_generation = 41
    with pytest.raises(native.KueueB200Error):
        ev.run_cycle(bad)


def test_usage_tracker_drives_incremental_cycles(ev):
    """api.UsageTracker (cache-side bookkeeping) + usage_delta_*: a sequence of cycles in which admitted workloads'
    usage is applied between cycles decides like the oracle on the full tables, including across a spec change."""
    from kueue_b200.api import UsageTracker
    t = UsageTracker()
    rng = np.random.default_rng(23)
    mk = lambda: synth.make_snapshot(3, W=2000, Q=200, heads="one_per_cq")  # noqa: E731
    usage = np.array(mk().arrays["cq_usage"]).reshape(200, -1).copy()
    gen = 5
    for cycle in range(5):
        if cycle == 3:
            gen = 6  # a ClusterQueue spec changed: static tables and the usage table travel again
        full = mk(); full.set("cq_usage", usage); full.static_generation = gen
        want = oracle.run_cycle(full)
        sent = mk(); sent.set("cq_usage", usage); sent.static_generation = gen
        sent = t.prepare(sent)
        if cycle not in (0, 3):
            assert "usage_delta_cq" in sent.arrays and not (sent.flags & abi.F_USAGE_RESIDENT)
        assert_cycle_equal(ev.run_cycle(sent, abi.CycleOut(full)), want)
        # the cache applies the admissions (cache.AssumeWorkload) and some workloads finish
        adm = np.flatnonzero(want.decision == 5)
        cqs = np.asarray(full.wl_cq)[np.asarray(full.heads)[adm]]
        usage = np.array(want.node_usage).reshape(-1, full.n_fr)[:200].copy()
        for c in cqs:
            t.touch(int(c))
        done = rng.choice(200, 7, replace=False)
        usage[done] = (usage[done] * 0.8).astype(np.int64)   # not touched: the tracker's diff must catch them
