"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "3", "--warmup", "1"],
                         check=True, capture_output=True, text=True, cwd=ROOT).stdout.strip().splitlines()
    line = json.loads(out[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_algorithmic_bytes_table_covers_every_timed_kernel():
    sys.path.insert(0, ROOT)
    import bench
    from kueue_b200 import abi, synth
    ab = bench.algorithmic_bytes(synth.make_snapshot(1))
    for name in abi.KERNEL_NAMES:
        # the kb_tas_find kernels are accounted in bench.run_tas (cfg5), "-" is an unused timing slot
        assert name in ab or name in ("k_lone", "k_tas_leaf", "k_tas_reduce", "k_tas_select", "-"), name
