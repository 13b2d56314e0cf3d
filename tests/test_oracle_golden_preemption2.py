"""Pins the oracle to TestHierarchicalPreemptions (preemption_hierarchical_test.go:46, 15 cases)
and TestFairPreemptions (preemption_fair_test.go:46, 31 cases); fixtures by tools/transcribe_tables.py."""
import json
import os

import pytest

from kueue_b200 import abi
from tests.test_oracle_golden_preemption import check_targets, run_case

HERE = os.path.dirname(__file__)
HIER = json.load(open(os.path.join(HERE, "golden", "preemption_hierarchical_cases.json")))
FAIR = json.load(open(os.path.join(HERE, "golden", "preemption_fair_cases.json")))


@pytest.mark.parametrize("name", list(HIER))
def test_hierarchical_preemption(name):
    tc = HIER[name]
    check_targets(run_case(tc), tc)


def fair_flags(tc):
    f = abi.FLAGS_DEFAULT | abi.F_FAIR_SHARING
    st = tc.get("strategies") or []
    if st:  # parseStrategies preemption.go:319-333
        f &= ~(abi.F_FS_STRATEGY_S2A | abi.F_FS_STRATEGY_S2B)
        for i, name in enumerate(st):
            f |= abi.F_FS_STRATEGY_S2A if name == "LessThanOrEqualToFinalShare" else abi.F_FS_STRATEGY_S2B
        if st[0] == "LessThanInitialShare" and len(st) > 1:
            f |= abi.F_FS_STRATEGY_S2B_FIRST
    return f


@pytest.mark.parametrize("name", list(FAIR))
def test_fair_preemption(name):
    tc = FAIR[name]
    assert run_case(tc, fair_flags(tc)) == tc["want"], tc["source"]
