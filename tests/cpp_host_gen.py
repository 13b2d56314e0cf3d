"""Generates a C++ translation unit that rebuilds scheduling scenarios with the C++ host mirror
(include/kueue_b200_host.hpp) from the same builder objects the Python tests use, so the two host sides can be
compared array by array (CPU) and decision by decision (GPU)."""
import os
import subprocess

from kueue_b200 import abi
from kueue_b200.api import queue_order_timestamp, resource_value

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _s(x):
    return '"' + str(x).replace("\\", "\\\\").replace('"', '\\"') + '"'


def _opt(name, q):
    return "std::nullopt" if q is None else f"(int64_t){resource_value(name, q)}LL"


def _rgs(holder):
    out = []
    for rg in holder.resource_groups:
        fqs = []
        for fq in rg:
            rs = ", ".join(f"kb::ResourceQuota{{{_s(r)}, (int64_t){resource_value(r, q.nominal)}LL, {_opt(r, q.borrowing_limit)}, {_opt(r, q.lending_limit)}}}"
                           for r, q in fq.resources.items())
            fqs.append(f"kb::FlavorQuotas{{{_s(fq.name)}, {{{rs}}}}}")
        out.append("{" + ", ".join(fqs) + "}")
    return "{" + ", ".join(out) + "}"


def _workload(w, flags, admitted):
    lines = ["  { kb::WorkloadInfo w;", f"    w.key = {_s(w.name)}; w.clusterQueue = {_s(w.admission.cq if admitted else w.cq)}; w.priority = {w.priority}; w.uid = {w.uid};"]
    if admitted:
        lines.append(f"    w.queueOrderTimestampNs = {w.creation_ns}LL; w.evicted = {'true' if w.evicted else 'false'};")
        if w.quota_reserved_ns is not None:
            lines.append(f"    w.quotaReservedNs = {w.quota_reserved_ns}LL;")
        for psa in w.admission.podsets:
            for r, (f, q) in psa.items():
                lines.append(f"    w.usage.push_back({{{_s(f)}, {_s(r)}, (int64_t){resource_value(r, q)}LL}});")
        lines.append("    s.admitted.push_back(w); }")
    else:
        ts = queue_order_timestamp(w, "Eviction", bool(flags & abi.F_PRIORITY_SORTING_WITHIN_COHORT))
        lines.append(f"    w.queueOrderTimestampNs = {ts}LL; w.lastAssignmentGeneration = {w.last_gen if w.last_tried is not None else -1};")
        for pi, ps in enumerate(w.podsets):
            lines.append(f"    {{ kb::PodSet p; p.name = {_s(ps.name)}; p.count = {ps.count};")
            if ps.min_count is not None:
                lines.append(f"      p.minCount = {ps.min_count};")
            for r, q in ps.requests.items():
                lines.append(f"      p.requests[{_s(r)}] = (int64_t){resource_value(r, q)}LL;")
            if ps.flavor_ok is not None:
                lines.append("      p.eligibleFlavors = std::vector<std::string>{" + ", ".join(_s(f) for f in ps.flavor_ok) + "};")
            if w.last_tried is not None and pi < len(w.last_tried):
                for r, v in w.last_tried[pi].items():
                    lines.append(f"      p.lastTriedFlavorIdx[{_s(r)}] = {v};")
            lines.append("      w.podSets.push_back(p); }")
        lines.append("    heads.push_back(w); }")
    return "\n".join(lines)


def scenario_cpp(k, cqs, cohorts, pending, admitted, flags, now_ns, flavors):
    L = [f"static void scenario_{k}(kb::Snapshot &s, std::vector<kb::WorkloadInfo> &heads, uint32_t &flags, int64_t &now) {{",
         f"  flags = {flags}u; now = {now_ns}LL;",
         "  s.resourceFlavors = {" + ", ".join(_s(f) for f in flavors) + "};"]
    for c in cqs:
        L.append(f"  {{ kb::ClusterQueue c; c.name = {_s(c.name)}; c.cohort = {_s(c.cohort or '')}; c.resourceGroups = {_rgs(c)};")
        L.append(f"    c.withinClusterQueue = {c.within_cluster_queue}; c.reclaimWithinCohort = {c.reclaim_within_cohort}; c.borrowWithinCohort = {c.borrow_within_cohort};")
        if c.bwc_threshold is not None:
            L.append(f"    c.maxPriorityThreshold = {c.bwc_threshold};")
        L.append(f"    c.whenCanBorrow = {c.when_can_borrow}; c.whenCanPreempt = {c.when_can_preempt}; c.preference = {c.preference}; c.queueingStrategy = {c.strategy};")
        L.append(f"    c.fairWeight = {c.fair_weight!r}; c.allocatableResourceGeneration = {c.generation}; s.clusterQueues.push_back(c); }}")
    for c in cohorts:
        L.append(f"  {{ kb::Cohort c; c.name = {_s(c.name)}; c.parent = {_s(c.parent or '')}; c.resourceGroups = {_rgs(c)}; c.fairWeight = {c.fair_weight!r}; s.cohorts.push_back(c); }}")
    for w in admitted:
        L.append(_workload(w, flags, True))
    for w in pending:
        L.append(_workload(w, flags, False))
    L.append("}")
    return "\n".join(L)


def build_driver(scenarios, workdir):
    """scenarios: list of scenario_cpp(...) strings.  Returns the path of the compiled driver."""
    os.makedirs(workdir, exist_ok=True)
    src = os.path.join(workdir, "scenarios.inc")
    with open(src, "w") as f:
        f.write("\n\n".join(scenarios))
        f.write("\n\ntypedef void (*scenario_fn)(kb::Snapshot &, std::vector<kb::WorkloadInfo> &, uint32_t &, int64_t &);\n")
        f.write("static scenario_fn SCENARIOS[] = {" + ", ".join(f"scenario_{k}" for k in range(len(scenarios))) + "};\n")
    exe = os.path.join(workdir, "host_driver")
    lib = os.path.join(ROOT, "kueue_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", workdir,
                           os.path.join(ROOT, "tests", "cpp", "host_driver.cpp"), "-o", exe,
                           "-L", lib, "-lkueue_b200", f"-Wl,-rpath,{lib}"])
    return exe
