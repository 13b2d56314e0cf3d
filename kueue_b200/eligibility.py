"""Host-side flavor eligibility — what the Go shim evaluates per (podset, flavor) before the
device pass and ships as the `ps_flavor_ok` bitmask (include/kueue_b200.h).

Restates checkFlavorForPodSets / flavorSelector
(pkg/scheduler/flavorassigner/flavorassigner.go:899-944, 965-1009): untolerated
NoSchedule/NoExecute taints of the flavor (corev1helpers.FindMatchingUntoleratedTaint,
k8s.io/component-helpers v0.35.2) and the podset's required node affinity / nodeSelector,
restricted to the resource group's label keys, matched against the flavor's node labels
(nodeaffinity.RequiredNodeAffinity.Match).  TAS checks are out of scope.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional


def tolerates(tol: dict, taint: dict) -> bool:
    """corev1.Toleration.ToleratesTaint."""
    if tol.get("effect") and tol["effect"] != taint.get("effect"):
        return False
    if tol.get("key") and tol["key"] != taint.get("key"):
        return False
    op = tol.get("operator") or "Equal"
    if op == "Exists":
        return True
    if op == "Equal":
        return (tol.get("value") or "") == (taint.get("value") or "")
    return False


def untolerated_taint(taints: Iterable[dict], tolerations: List[dict]) -> Optional[dict]:
    for t in taints:
        if t.get("effect") not in ("NoSchedule", "NoExecute"):
            continue
        if not any(tolerates(tol, t) for tol in tolerations):
            return t
    return None


def _match_expr(e: dict, labels: Dict[str, str]) -> bool:
    k, op, vals = e["key"], e["operator"], e.get("values") or []
    has = k in labels
    if op == "In":
        return has and labels[k] in vals
    if op == "NotIn":
        return not (has and labels[k] in vals)
    if op == "Exists":
        return has
    if op == "DoesNotExist":
        return not has
    if op in ("Gt", "Lt"):
        try:
            a, b = int(labels[k]), int(vals[0])
        except (KeyError, ValueError, IndexError):
            return False
        return a > b if op == "Gt" else a < b
    return False


def affinity_matches(node_selector: Optional[Dict[str, str]], terms: Optional[List[dict]], allowed_keys: set,
                     labels: Dict[str, str]) -> bool:
    """flavorSelector(spec, allowedKeys).Match(node with `labels`)."""
    for k, v in (node_selector or {}).items():
        if k in allowed_keys and labels.get(k) != v:
            return False
    kept = []
    for t in terms or []:
        exprs = [e for e in (t.get("matchExpressions") or []) if e["key"] in allowed_keys]
        if not exprs:  # an empty term matches everything and terms are ORed
            kept = []
            break
        kept.append(exprs)
    if not kept:
        return True
    return any(all(_match_expr(e, labels) for e in exprs) for exprs in kept)


def flavor_eligible(podset: dict, flavor: dict, rg_label_keys: set) -> bool:
    """podset: {tolerations, nodeSelector, affinityTerms}; flavor: {nodeLabels, taints, tolerations}."""
    tols = list(podset.get("tolerations") or []) + list(flavor.get("tolerations") or [])
    if untolerated_taint(flavor.get("taints") or [], tols) is not None:
        return False
    return affinity_matches(podset.get("nodeSelector"), podset.get("affinityTerms"), rg_label_keys, flavor.get("nodeLabels") or {})
