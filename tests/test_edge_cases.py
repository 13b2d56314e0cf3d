"""Empty and ragged inputs: the oracle accepts them (CPU) and the device agrees (GPU)."""
import numpy as np
import pytest

import oracle
from kueue_b200 import abi
from tests.edge_cases import edge_snapshots
from tests.helpers import assert_cycle_equal

SNAPS = edge_snapshots()


@pytest.mark.parametrize("name", list(SNAPS))
def test_oracle_edge(name):
    snap = SNAPS[name]
    out = oracle.run_cycle(snap, 64)
    assert len(out.decision) == snap.n_heads
    if name == "single workload":
        assert out.decision.tolist() == [abi.DEC_ASSUMED]
    if name in ("uncovered resource and oversize", "oversize request"):
        assert out.decision.tolist() == [abi.DEC_NOFIT]
    if name == "lone queue preempts all":
        assert out.decision.tolist() == [abi.DEC_PREEMPTING] and out.n_targets == 4
    if name == "ragged podsets":
        assert (out.decision == abi.DEC_ASSUMED).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SNAPS))
def test_device_edge(name):
    from kueue_b200 import native
    ev = native.Evaluator(0)
    try:
        snap = SNAPS[name]
        assert_cycle_equal(ev.run_cycle(snap, abi.CycleOut(snap, 64)), oracle.run_cycle(snap, 64))
    finally:
        ev.close()
