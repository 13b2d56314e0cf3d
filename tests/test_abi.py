"""CPU-side checks of the boundary: the shared library loads, exports every symbol the
header declares, and the ctypes structs have the C layout (no compute without a GPU)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

from kueue_b200 import abi, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "kueue_b200.h")


def test_exports_every_declared_symbol():
    text = open(HDR).read()
    declared = set(re.findall(r"\b(kb_[a-z_]+)\s*\(", text))
    declared -= {"kb_handle"}
    lib = native.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(native.EXPORTS) <= declared


def test_struct_layout_matches_c():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "kueue_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(kb_snapshot), sizeof(kb_cycle_out), sizeof(kb_tree_out), sizeof(kb_stats), sizeof(kb_config));
  printf("%zu %zu %zu %zu\n", offsetof(kb_snapshot, now_ns), offsetof(kb_snapshot, parent), offsetof(kb_snapshot, heads), offsetof(kb_cycle_out, node_usage));
  printf("%zu %zu %zu %zu\n", offsetof(kb_snapshot, static_generation), offsetof(kb_snapshot, n_usage_delta), offsetof(kb_snapshot, usage_delta_cq), offsetof(kb_snapshot, usage_delta_rows));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "probe")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe], text=True).split()
    sizes = [int(x) for x in out]
    assert sizes[:5] == [C.sizeof(abi.kb_snapshot), C.sizeof(abi.kb_cycle_out), C.sizeof(abi.kb_tree_out),
                         C.sizeof(abi.kb_stats), C.sizeof(abi.kb_config)]
    assert sizes[5:9] == [abi.kb_snapshot.now_ns.offset, abi.kb_snapshot.parent.offset, abi.kb_snapshot.heads.offset,
                          abi.kb_cycle_out.node_usage.offset]
    assert sizes[9:] == [abi.kb_snapshot.static_generation.offset, abi.kb_snapshot.n_usage_delta.offset,
                         abi.kb_snapshot.usage_delta_cq.offset, abi.kb_snapshot.usage_delta_rows.offset]


def test_no_device_is_a_loud_error():
    """Without a GPU the product path must fail, never fall back to the CPU."""
    import pytest
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(native.KueueB200Error):
        native.Evaluator(0)
