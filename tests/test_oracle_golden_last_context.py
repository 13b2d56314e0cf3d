"""Pins the multi-cycle LastAssignment handling (LastTriedFlavorIdx fed back as ps_last_tried) to
TestLastSchedulingContext (pkg/scheduler/scheduler_test.go:8569)."""
import pytest

import oracle
from tests.last_context_golden import DOC, check


@pytest.mark.parametrize("name", list(DOC["cases"]))
def test_last_scheduling_context(name):
    check(DOC["cases"][name], oracle.run_cycle)
