"""Builds kueue_b200.tas objects from the TestFindTopologyAssignments fixtures (tests/golden/tas_cases.json)."""
import json
import os

from kueue_b200 import abi, tas

DOC = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tas_cases.json")))


def build(tc):
    non_tas = {}
    for p in tc["pods"]:
        if p["phase"] in ("Failed", "Succeeded") or not p["node"]:
            continue  # utilpod.IsTerminated / unscheduled pods use no node capacity (tas_non_tas_pod_cache.go:45-82)
        u = non_tas.setdefault(p["node"], {})
        for r, q in p["requests"].items():
            u[r] = u.get(r, 0) + q
        u["pods"] = u.get("pods", 0) + 1
    res = sorted({r for ps in tc["podSets"] for r in ps["requests"]})
    # non-TAS usage is keyed by node name; the fixtures name nodes "<block>-<rack>-<host>", the hostname label is what the pod binds to
    nodes = []
    for n in tc["nodes"]:
        n = dict(n)
        nodes.append(n)
    by_host = {n["labels"].get(tas.HOSTNAME, n["name"]): n["name"] for n in nodes}
    non_tas_by_name = {}
    for node, u in non_tas.items():
        non_tas_by_name[by_host.get(node, node)] = u
    topo = tas.TasTopology(tc["levels"], nodes, node_labels=tc["nodeLabels"], resources=res, non_tas_usage=non_tas_by_name)
    reqs = tas.TasRequests(topo)
    mixed = True  # features.TASProfileMixed defaults to on
    for ps in tc["podSets"]:
        reqs.add(0, ps["requests"], ps["count"], ps["topologyRequest"], tolerations=ps["tolerations"], node_selector=ps["nodeSelector"],
                 profile_mixed=mixed)
    return topo, reqs.finalize()


def check(tc, out, topo):
    for q, ps in enumerate(tc["podSets"]):
        want = ps["wantAssignment"]
        if want is None:
            assert out.status[q] != abi.TAS_OK, (q, "expected failure: " + ps["wantReason"])
            break  # the reference returns at the first failure
        assert out.status[q] == abi.TAS_OK, (q, int(out.status[q]))
        got = [(topo.leaf_values[leaf], cnt) for leaf, cnt in out.assignment(q)]
        nl = len(want["levels"])
        exp = [(tuple(d["values"]), d["count"]) for d in want["domains"]]
        assert [(v[len(v) - nl:], c) for v, c in got] == exp, (q, got, exp)
