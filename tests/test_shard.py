"""N>1 path on CPU: sharding by root cohort + merge reproduce the single-shard result.
The per-shard evaluator here is the oracle (no GPU in this container); on the GPU box
bench.py --gpus N drives the same shard/merge code with the CUDA evaluator."""
import os
import socket

import numpy as np
import pytest

import oracle
from kueue_b200 import abi, shard, synth
from tests.helpers import assert_cycle_equal

CASES = [
    lambda: synth.make_snapshot(2, W=3000, Q=60),
    lambda: synth.make_snapshot(3, W=3000, Q=300, heads="one_per_cq"),
    lambda: synth.make_snapshot(4, W=40000, Q=2000, heads="one_per_cq"),
    lambda: synth.make_snapshot(2, W=3000, Q=30, preemption=True, tight=1.2),
]


@pytest.mark.parametrize("make", CASES)
@pytest.mark.parametrize("world", [2, 3])
def test_shard_merge_equals_full(make, world):
    snap = make()
    cap = 40 * snap.n_adm + 10000
    want = oracle.run_cycle(snap, cap)
    node_rank = shard.partition_roots(snap, world)
    assert len(np.unique(node_rank)) == min(world, int((snap.arrays["parent"] < 0).sum()))
    parts = []
    for r in range(world):
        sub, m = shard.shard(snap, r, world, node_rank)
        parts.append((oracle.run_cycle(sub, cap), m))
    assert_cycle_equal(shard.merge(snap, parts), want)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    snap = synth.make_snapshot(4, W=40000, Q=2000, heads="one_per_cq")
    cap = 40 * snap.n_adm + 10000
    sub, m = shard.shard(snap, rank, world)
    out = oracle.run_cycle(sub, cap)
    payload = {f: getattr(out, f) for f in ("decision", "mode", "borrow", "commit_rank", "ps_flavor", "ps_res_mode",
                                            "ps_tried_idx", "ps_count", "tgt_start", "tgt_adm", "tgt_reason", "node_usage")}
    gathered = [None] * world
    dist.all_gather_object(gathered, (payload, m))
    if rank == 0:
        class O:  # noqa: simple attribute bag
            pass
        parts = []
        for p, mm in gathered:
            o = O()
            for k, v in p.items():
                setattr(o, k, v)
            parts.append((o, mm))
        full = shard.merge(snap, parts)
        want = oracle.run_cycle(snap, cap)
        try:
            assert_cycle_equal(full, want)
            q.put("ok")
        except AssertionError as e:  # pragma: no cover
            q.put(f"mismatch: {e}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_over_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert res == "ok", res
