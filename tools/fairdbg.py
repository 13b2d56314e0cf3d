import json, os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import oracle
from kueue_b200 import native, abi
from tests.golden_loader import build_preemption_case
from tests.test_oracle_golden_preemption2 import fair_flags
from tests.helpers import assert_cycle_equal
cases = json.load(open('/root/repo/tests/golden/preemption_fair_cases.json'))
ev = native.Evaluator(0)
names = list(cases)
start = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for i, name in enumerate(names[start:], start):
    tc = cases[name]
    snap, idx = build_preemption_case(tc, fair_flags(tc))
    print(i, name, flush=True)
    want = oracle.run_cycle(snap)
    try:
        got = ev.run_cycle(snap)
        assert_cycle_equal(got, want)
        print('   ok', flush=True)
    except Exception as e:
        print('   FAIL', str(e)[:300], flush=True)
