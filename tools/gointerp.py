"""Evaluates the AST of tools/goparse.py for the builder vocabulary of
pkg/util/testing/v1beta2/wrappers.go into plain dicts (JSON-able fixtures).
Transcription tooling only — see goparse.py."""
from __future__ import annotations

import copy

NOW = 1_700_000_000_000_000_000  # `now := time.Now().Truncate(time.Second)` in the tests
DUR = {"Nanosecond": 1, "Microsecond": 10**3, "Millisecond": 10**6, "Second": 10**9, "Minute": 60 * 10**9, "Hour": 3600 * 10**9}
CONST = {
    "corev1.ResourceCPU": "cpu", "corev1.ResourceMemory": "memory", "corev1.ResourcePods": "pods",
    "corev1.ResourceEphemeralStorage": "ephemeral-storage", "corev1.LabelHostname": "kubernetes.io/hostname",
    "kueue.DefaultPodSetName": "main",
    "utiltesting.Ki": 2**10, "utiltesting.Mi": 2**20, "utiltesting.Gi": 2**30, "utiltesting.Ti": 2**40,
    "metav1.ConditionTrue": "True", "metav1.ConditionFalse": "False", "metav1.NamespaceDefault": "default",
}
IGNORED = set()


class Obj(dict):
    """A builder object; unknown methods are no-ops recorded in IGNORED."""

    def __init__(self, kind, **kw):
        super().__init__(kw)
        self["_kind"] = kind


def strip(v):
    if isinstance(v, Obj):
        return {k: strip(x) for k, x in v.items() if k != "_kind"} | {"_kind": v["_kind"]}
    if isinstance(v, dict):
        return {k: strip(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [strip(x) for x in v]
    return v


def _symtail(v):
    if isinstance(v, tuple) and v and v[0] == "sym":
        v = v[1].split(".")[-1]
        for pre in ("TaintEffect", "TolerationOp", "NodeSelectorOp"):
            if v.startswith(pre):
                return v[len(pre):]
    return v


def _k8s(d):
    """corev1.Taint / corev1.Toleration literal -> lower-case keyed dict with plain strings."""
    return {k[0].lower() + k[1:]: _symtail(v) for k, v in d.items() if not str(k).startswith("_")}


def _term(t):
    exprs = []
    for e in t.get("MatchExpressions") or []:
        exprs.append({"key": e.get("Key"), "operator": _symtail(e.get("Operator")), "values": list(e.get("Values") or [])})
    return {"matchExpressions": exprs}


class Interp:
    def __init__(self, env=None, helpers=None):
        self.env = dict(env or {})
        self.helpers = helpers or {}

    # ------------------------------------------------------------------
    def ev(self, n):
        k = n[0]
        if k == "str":
            return n[1]
        if k == "num":
            return n[1]
        if k == "id":
            name = n[1]
            if name in self.env:
                return self.env[name]
            if name in ("true", "false"):
                return name == "true"
            if name == "nil":
                return None
            return ("sym", name)
        if k == "type":
            return ("sym", n[1])
        if k == "func":
            return ("funclit", n[1]) if n[1] is not None else None
        if k == "unary":
            v = self.ev(n[2])
            if n[1] == "-":
                return -v
            return v
        if k == "bin":
            a, b = self.ev(n[2]), self.ev(n[3])
            return {"+": lambda: a + b, "-": lambda: a - b, "*": lambda: a * b, "/": lambda: a // b}[n[1]]()
        if k == "sel":
            q = self.qual(n)
            if q in CONST:
                return CONST[q]
            if q and q.startswith("time.") and q[5:] in DUR:
                return DUR[q[5:]]
            base = self.ev(n[1])
            if isinstance(base, Obj):
                if n[2] in ("Cohort", "ClusterQueue", "Workload", "PodSet"):  # embedded struct access, e.g. MakeCohort(..).Cohort
                    return base
                return ("method", base, n[2])
            if isinstance(base, dict) and n[2] in base:
                return base[n[2]]
            if isinstance(base, tuple) and base and base[0] == "sym":
                return ("sym", f"{base[1]}.{n[2]}")
            if isinstance(base, int) and n[2] in ("Add", "Truncate", "Sub"):
                return ("timefn", base, n[2])
            return ("sym", f"?.{n[2]}")
        if k == "index":
            base = self.ev(n[1])
            return base  # generic instantiation ptr.To[int32]
        if k == "comp":
            return self.comp(n)
        if k == "call":
            return self.call(n)
        raise ValueError(n)

    def qual(self, n):
        if n[0] == "id":
            return n[1]
        if n[0] == "sel":
            b = self.qual(n[1])
            return f"{b}.{n[2]}" if b else None
        return None

    def comp(self, n):
        typ = n[1][1] if n[1] else None
        elems = n[2]
        keyed = any(k is not None for k, _ in elems)
        if typ and typ.startswith("[]") or (typ is None and not keyed):
            return [self.ev(v) for _, v in elems]
        out = {}
        for k, v in elems:
            if k[0] == "id" and (k[1] not in self.env or not (typ or "").startswith("map[")):
                key = k[1]
            else:
                key = self.ev(k)
                if isinstance(key, tuple) and key[0] == "sym":
                    key = key[1]
                if isinstance(key, dict):
                    key = tuple(sorted((a, b) for a, b in key.items() if a != "_type"))
            out[key] = self.ev(v)
        if typ and not typ.startswith("map["):
            out["_type"] = typ
        return out

    # ------------------------------------------------------------------
    def call(self, n):
        fn, args = n[1], n[2]
        if fn[0] == "index":  # generic instantiation: ptr.To[int32](x)
            fn = fn[1]
        q = self.qual(fn)
        if q:
            short = q.split(".")[-1]
            if q in ("time.Now",):
                return NOW
            if short in self.helpers and fn[0] == "id":
                return self.helpers[short](self, [self.ev(a) for a in args])
            if q.startswith(("utiltestingapi.", "utiltesting.", "testingapi.", "testingnode.", "testingpod.")) and short.startswith("Make"):
                return self.make(short, [self.ev(a) for a in args])
            if q in ("resource.MustParse", "metav1.NewTime", "ptr.To", "kueue.ResourceFlavorReference", "kueue.ClusterQueueReference",
                     "kueue.CohortReference", "kueue.PodSetReference", "kueue.LocalQueueName", "corev1.ResourceName", "int32", "int64",
                     "types.UID", "workload.Reference"):
                vals = [self.ev(a) for a in args]
                return vals[0] if len(vals) == 1 else vals
            if q == "sets.New":
                return [self.ev(a) for a in args]
            if q == "utiltesting.SingleContainerForRequest":
                m = self.ev(args[0])
                return [{"_container": True, "requests": {k2: v2 for k2, v2 in m.items() if not str(k2).startswith("_")}}]
        if fn[0] == "sel" and not (q and q in CONST):
            base = self.ev(fn[1])
            if isinstance(base, Obj):
                return self.method(base, fn[2], [self.ev(a) for a in args])
        f = self.ev(fn)
        vals = [self.ev(a) for a in args]
        if isinstance(f, tuple) and f[0] == "funclit":  # immediately-invoked helper: run its assignments, return its value
            saved = dict(self.env)
            try:
                for kind, name, expr in f[1]:
                    if kind == "assign":
                        self.env[name] = self.ev(expr)
                    else:
                        return self.ev(expr)
                return None
            finally:
                self.env = saved
        if isinstance(f, tuple) and f[0] == "method":
            return self.method(f[1], f[2], vals)
        if isinstance(f, tuple) and f[0] == "timefn":
            if f[2] == "Add":
                return f[1] + vals[0]
            return f[1]
        if isinstance(f, tuple) and f[0] == "sym":
            if f[1].endswith((".Truncate", ".Obj", ".DeepCopy")):
                return NOW
            return {"_call": f[1], "args": vals}
        if isinstance(f, int):  # ptr.To[int32](x) evaluated through index -> call on the value
            return vals[0] if vals else f
        return {"_call": str(f), "args": vals}

    def make(self, short, a):
        if short == "MakeClusterQueue":
            return Obj("ClusterQueue", name=a[0], cohort=None, resourceGroups=[], preemption={}, fairWeight=None,
                       flavorFungibility={}, queueingStrategy=None)
        if short == "MakeCohort":
            return Obj("Cohort", name=a[0], parent=None, resourceGroups=[], fairWeight=None)
        if short == "MakeFlavorQuotas":
            return Obj("FlavorQuotas", flavor=a[0], resources=[])
        if short == "MakeWorkload":
            return Obj("Workload", name=a[0], ns=a[1] if len(a) > 1 else "", priority=0, creation=NOW, uid=None,
                       podsets=[{"name": "main", "count": 1, "minCount": None, "requests": {}}], admission=None,
                       reservedAt=None, conditions=[], queue=None)
        if short == "MakePodSet":
            return Obj("PodSet", name=a[0], count=a[1], minCount=None, requests={}, tolerations=[], nodeSelector=None, affinityTerms=None)
        if short == "MakeAdmission":
            return Obj("Admission", cq=a[0], podsets=[{"name": nm, "count": 1, "assignments": {}} for nm in (a[1:] or ["main"])])
        if short == "MakePodSetAssignment":
            return Obj("PodSetAssignment", name=a[0], count=1, assignments={})
        if short == "MakeResourceFlavor":
            return Obj("ResourceFlavor", name=a[0], nodeLabels={}, taints=[], tolerations=[])
        if short == "MakeLocalQueue":
            return Obj("LocalQueue", name=a[0], ns=a[1], cq=None)
        return Obj(short, args=a)

    def method(self, o, m, a):
        k = o["_kind"]
        if m in ("Obj", "DeepCopy"):
            return o
        if m == "Clone":
            return copy.deepcopy(o)
        if k in ("MakeNode", "MakePod"):  # testingnode / testingpod wrappers: keep the call chain, tools/transcribe_tas.py reads it
            o.setdefault("calls", []).append([m, [strip(x) for x in a]])
            return o
        if k == "FlavorQuotas":
            if m == "Resource":
                o["resources"].append({"name": a[0], "nominal": a[1] if len(a) > 1 else "0",
                                       "borrowingLimit": a[2] if len(a) > 2 and a[2] != "" else None,
                                       "lendingLimit": a[3] if len(a) > 3 and a[3] != "" else None})
                return o
            if m == "ResourceQuotaWrapper":
                return Obj("RQW", parent=o, q={"name": a[0], "nominal": "0", "borrowingLimit": None, "lendingLimit": None})
        if k == "RQW":
            if m == "NominalQuota": o["q"]["nominal"] = a[0]; return o
            if m == "BorrowingLimit": o["q"]["borrowingLimit"] = a[0]; return o
            if m == "LendingLimit": o["q"]["lendingLimit"] = a[0]; return o
            if m == "Append": o["parent"]["resources"].append(o["q"]); return o["parent"]
        if k in ("ClusterQueue", "Cohort"):
            if m == "ResourceGroup":
                if a:
                    o["resourceGroups"].append([strip(x) for x in a])
                return o
            if m == "Cohort": o["cohort"] = a[0]; return o
            if m == "Parent": o["parent"] = a[0]; return o
            if m == "FairWeight": o["fairWeight"] = a[0]; return o
            if m == "Preemption": o["preemption"] = a[0]; return o
            if m == "FlavorFungibility": o["flavorFungibility"] = a[0]; return o
            if m == "QueueingStrategy": o["queueingStrategy"] = a[0]; return o
            if m == "NamespaceSelector": o["namespaceSelector"] = a[0]; return o
        if k == "LocalQueue":
            if m == "ClusterQueue": o["cq"] = a[0]; return o
        if k == "Workload":
            if m == "Priority": o["priority"] = a[0]; return o
            if m == "Name": o["name"] = a[0]; return o
            if m == "Creation": o["creation"] = a[0]; return o
            if m == "UID": o["uid"] = a[0]; return o
            if m == "Queue": o["queue"] = a[0]; return o
            if m == "Request": o["podsets"][0]["requests"][a[0]] = a[1]; return o
            if m == "PodSets": o["podsets"] = [strip(x) for x in a]; return o
            if m in ("ReserveQuota", "ReserveQuotaAt"):  # replaces Status.Conditions (wrappers.go:167-177)
                o["admission"] = strip(a[0]); o["reservedAt"] = a[1] if len(a) > 1 else NOW; o["conditions"] = []; return o
            if m in ("Condition", "SetOrReplaceCondition"): o["conditions"].append(a[0]); return o
            if m == "Admitted": return o
            if m == "Admission": o["admission"] = strip(a[0]) if a[0] is not None else None; return o
            if m == "SimpleReserveQuota":  # wrappers.go:150-164: every request of podset 0 on one flavor, x Count
                ps = o["podsets"][0]
                o["admission"] = {"cq": a[0], "_kind": "Admission", "podsets": [{
                    "name": ps["name"], "count": ps["count"],
                    "assignments": {r: [a[1], q, ps["count"]] for r, q in ps["requests"].items()}}]}
                o["reservedAt"] = a[2] if len(a) > 2 else NOW
                o["conditions"] = []
                return o
            if m == "QuotaReservedTime": o["reservedAt"] = a[0]; return o
        if k == "PodSet":
            if m == "Request": o["requests"][a[0]] = a[1]; return o
            if m == "Limit": o.setdefault("limits", {})[a[0]] = a[1]; return o
            if m == "SetMinimumCount": o["minCount"] = a[0]; return o
            if m == "PodSetGroup": o["group"] = a[0]; return o
            if m == "Containers":
                cs = [c for x in a for c in (x if isinstance(x, list) else [x])]
                o["requests"] = dict(cs[0]["requests"]) if cs else {}
                return o
            if m == "Toleration": o["tolerations"].append(_k8s(a[0])); return o
            if m == "NodeSelector": o["nodeSelector"] = {k2: v2 for k2, v2 in a[0].items() if not str(k2).startswith("_")}; return o
            if m == "RequiredDuringSchedulingIgnoredDuringExecution":
                o["affinityTerms"] = (o["affinityTerms"] or []) + [_term(t) for t in a[0]]; return o
        if k == "ResourceFlavor":
            if m == "NodeLabel": o["nodeLabels"][a[0]] = a[1]; return o
            if m == "Taint": o["taints"].append(_k8s(a[0])); return o
            if m == "Toleration": o["tolerations"].append(_k8s(a[0])); return o
        if k == "Admission":
            if m == "PodSets": o["podsets"] = [strip(x) for x in a]; return o
            if m == "Assignment": o["podsets"][0]["assignments"][a[0]] = [a[1], a[2]]; return o
            if m == "AssignmentPodCount": o["podsets"][0]["count"] = a[0]; return o
            if m == "AssignmentWithIndex": o["podsets"][a[0]]["assignments"][a[1]] = [a[2], a[3]]; return o
            if m == "AssignmentPodCountWithIndex": o["podsets"][a[0]]["count"] = a[1]; return o
        if k == "PodSetAssignment":
            if m == "Assignment": o["assignments"][a[0]] = [a[1], a[2]]; return o
            if m == "Count": o["count"] = a[0]; return o
            if m == "Flavor": o["assignments"].setdefault(a[0], [a[1], "0"])[0] = a[1]; return o
            if m == "ResourceUsage": o["assignments"].setdefault(a[0], [None, a[1]])[1] = a[1]; return o
        IGNORED.add(f"{k}.{m}")
        return o
