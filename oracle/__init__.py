"""CPU parity oracle loader — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  kueue_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libkueue_oracle.so")
    src = os.path.join(_HERE, "kueue_oracle.cpp")
    src2 = os.path.join(_HERE, "kueue_oracle_tas.cpp")
    hdr = os.path.join(_HERE, "..", "include", "kueue_b200.h")
    stale = (not os.path.exists(so)) or any(os.path.getmtime(so) < os.path.getmtime(p) for p in (src, src2, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libkueue_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ko_tree_eval.restype = C.c_int32
        _LIB.ko_run_cycle.restype = C.c_int32
        _LIB.ko_get_targets.restype = C.c_int32
    return _LIB


def tree_eval(snap, wl_req=None):
    import numpy as np
    from kueue_b200 import abi
    out = abi.TreeOut(snap)
    s = snap.as_struct()
    if wl_req is not None:
        wl_req = np.ascontiguousarray(wl_req, np.int64)
        rc = lib().ko_tree_eval_req(C.byref(s), wl_req.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(out.struct))
    else:
        rc = lib().ko_tree_eval(C.byref(s), C.byref(out.struct))
    assert rc == 0
    return out


def run_cycle(snap, tgt_capacity=None):
    from kueue_b200 import abi
    out = abi.CycleOut(snap, tgt_capacity)
    s = snap.as_struct()
    rc = lib().ko_run_cycle(C.byref(s), C.byref(out.struct))
    assert rc == 0, rc
    out.n_targets = out.struct.n_targets
    return out


def get_targets(snap, wl, ps_flavor, ps_res_mode, ps_count=None, cap=4096):
    """Preemptor.GetTargets for workload `wl` with an explicit assignment (TestPreemption-style)."""
    import numpy as np
    s = snap.as_struct()
    f = np.ascontiguousarray(ps_flavor, np.int8); m = np.ascontiguousarray(ps_res_mode, np.int8)
    cnt = None if ps_count is None else np.ascontiguousarray(ps_count, np.int32)
    adm = np.zeros(cap, np.int32); reason = np.zeros(cap, np.uint8)
    n = lib().ko_get_targets(C.byref(s), C.c_int32(wl), f.ctypes.data_as(C.POINTER(C.c_int8)), m.ctypes.data_as(C.POINTER(C.c_int8)),
                             None if cnt is None else cnt.ctypes.data_as(C.POINTER(C.c_int32)),
                             adm.ctypes.data_as(C.POINTER(C.c_int32)), reason.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int32(cap))
    return [(int(adm[i]), int(reason[i])) for i in range(n)]


def assign_stub(snap, wl, stub_mode=None, stub_borrow=None):
    """FlavorAssigner.Assign with the reference's testOracle stub; returns a dict."""
    import numpy as np
    s = snap.as_struct()
    R, FR = snap.n_resource, snap.n_fr
    np_ = int(snap.wl_ps_start[wl + 1] - snap.wl_ps_start[wl])
    sm = np.full(FR, -1, np.int8) if stub_mode is None else np.ascontiguousarray(stub_mode, np.int8)
    sb = np.zeros(FR, np.int32) if stub_borrow is None else np.ascontiguousarray(stub_borrow, np.int32)
    fl = np.full((np_, R), -1, np.int8); md = np.full((np_, R), -1, np.int8); tr = np.full((np_, R), -1, np.int8)
    bor = C.c_int32(0); use = np.zeros(FR, np.int64)
    P8, P32, P64 = C.POINTER(C.c_int8), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    lib().ko_assign_stub.restype = C.c_int32
    mode = lib().ko_assign_stub(C.byref(s), C.c_int32(wl), sm.ctypes.data_as(P8), sb.ctypes.data_as(P32), fl.ctypes.data_as(P8),
                                md.ctypes.data_as(P8), tr.ctypes.data_as(P8), C.byref(bor), use.ctypes.data_as(P64))
    return {"mode": mode, "flavor": fl, "res_mode": md, "tried": tr, "borrowing": bor.value, "usage": use}


def resources_to_reserve(snap, cq, mode, borrowing, usage):
    """resourcesToReserve (scheduler.go:530-548): usage / result are [F*R] arrays, -1 = absent cell."""
    import numpy as np
    s = snap.as_struct()
    u = np.ascontiguousarray(usage, np.int64)
    out = np.full(snap.n_fr, -1, np.int64)
    rc = lib().ko_resources_to_reserve(C.byref(s), C.c_int32(cq), C.c_int32(mode), C.c_int32(borrowing),
                                       u.ctypes.data_as(C.POINTER(C.c_int64)), out.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0, rc
    return out


def entry_less(flags, a, b):
    """entryComparer.less (fair_sharing_iterator.go:166-199); a, b = (borrowing, priority, ts_ns, drs_ratio, drs_weight)."""
    f = lib().ko_entry_less
    f.restype = C.c_int32
    f.argtypes = [C.c_uint32] + [C.c_int32, C.c_int32, C.c_int64, C.c_double, C.c_double] * 2
    return bool(f(flags, *a, *b))


def satisfies_policy(snap, pre_prio, pre_ts, adm, policy):
    s = snap.as_struct()
    return bool(lib().ko_satisfies_policy(C.byref(s), C.c_int32(pre_prio), C.c_int64(pre_ts), C.c_int32(adm), C.c_int32(policy)))


def sort_candidates(snap, cq):
    import numpy as np
    s = snap.as_struct()
    order = np.zeros(snap.n_adm, np.int32)
    rc = lib().ko_sort_candidates(C.byref(s), C.c_int32(cq), order.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    return order.tolist()


def is_preferred(a, b, pref):
    """isPreferred (flavorassigner.go:410-441); a, b = (preemption mode 0..4, borrowing level)."""
    return bool(lib().ko_is_preferred(C.c_int32(a[0]), C.c_int32(a[1]), C.c_int32(b[0]), C.c_int32(b[1]), C.c_int32(pref)))


def tas_find(topo, reqs, capacity=None):
    """FindTopologyAssignmentsForFlavor restated (oracle/kueue_oracle_tas.cpp): topo / reqs are kueue_b200.tas objects."""
    from kueue_b200 import tas
    out = tas.TasOut(reqs, capacity if capacity is not None else max(16, int(reqs.count.sum()) + 16))
    f = lib().ko_tas_find
    f.restype = C.c_int32
    rc = f(C.byref(topo.struct), C.byref(reqs.struct), C.byref(out.struct))
    assert rc == 0, rc
    return out
