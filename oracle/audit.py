"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the audit sweep's aggregation
(SURVEY.md rows a-1..a-4): sequential Review of every object, per-constraint totals, the K-smallest
status violations under SVQueue ordering, 256-byte message truncation.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import heapq

from . import k8s

MSG_SIZE = 256                       # pkg/audit/manager.go:48 msgSize
DEFAULT_VIOLATIONS_LIMIT = 20        # pkg/audit/manager.go:64 --constraint-violations-limit


def truncate_string(s: str, size: int = MSG_SIZE) -> str:
    """truncateString -- pkg/audit/manager.go:1043-1052 (operates on BYTES: Go string slicing)."""
    b = s.encode("utf-8")
    if len(b) > size:
        if size > 3:
            size -= 3
        return b[:size].decode("utf-8", errors="surrogateescape") + "..."
    return s


def sv_key(v):
    """SVQueue.Less ordering key -- pkg/audit/manager.go:117-137: group, version, kind, namespace, name,
    message, enforcementAction (Go string compare = bytewise)."""
    return tuple(x.encode("utf-8") for x in (v["group"], v["version"], v["kind"], v["namespace"], v["name"],
                                             v["message"], v["enforcementAction"]))


class _Rev:
    __slots__ = ("k", "v")

    def __init__(self, v):
        self.k, self.v = sv_key(v), v

    def __lt__(self, o):  # max-heap on the key
        return self.k > o.k


class LimitQueue:
    """LimitQueue -- pkg/audit/manager.go:161-202: keeps the `limit` smallest violations."""

    def __init__(self, limit):
        self.limit, self.h = limit, []

    def push(self, v):
        heapq.heappush(self.h, _Rev(v))
        while len(self.h) > self.limit:
            heapq.heappop(self.h)

    def drain_descending(self):
        """updateConstraintStatus pops the max-heap: status.violations comes out in DESCENDING key order
        -- pkg/audit/manager.go:984-996."""
        out = []
        while self.h:
            out.append(heapq.heappop(self.h).v)
        return out


def audit(client: k8s.Client, objects, namespaces=None, excluded_namespaces=(), limit=DEFAULT_VIOLATIONS_LIMIT,
          source="Original", expansion=None):
    """reviewObjects + addAuditResponsesToUpdateLists -- pkg/audit/manager.go:668-777,886-945.

    objects: iterable of object dicts.  namespaces: {name: namespace object} (the audit's nsCache,
    :697-706).  expansion: an oracle.expansion.System -- every object is expanded and its resultants reviewed (:733-765).
    Returns {"totals": {(kind,name): n}, "by_action": {action: n},
    "violations": {(kind,name): [status violation dicts, descending]}, "results": [...], "expand_errors": {index: text}}."""
    from oracle import expansion as X
    namespaces = namespaces or {}
    totals, by_action, queues, all_results, expand_errors = {}, {}, {}, [], {}
    for idx, obj in enumerate(objects):
        if k8s.is_namespace_excluded(excluded_namespaces, obj):      # :531-538 skipExcludedNamespace
            continue
        ns_name = k8s._meta(obj, "namespace")
        ns = namespaces.get(ns_name) if ns_name else None              # :697-706
        review = k8s.Review(obj=obj, ns=ns, source=source)             # :707-711 AugmentedUnstructured
        results = list(client.review(review, k8s.AUDIT_EP))            # :720
        if expansion is not None:
            # :733-765 -- Expand fails: the error is logged and the loop `continue`s, past the bookkeeping of the object's own results
            try:
                resultants = expansion.expand(obj, (ns.get("metadata") or {}).get("name", "") if ns is not None else None)
            except X.ExpansionError as e:
                expand_errors[idx] = "unable to expand object: " + str(e)
                continue
            for robj, tname, action in resultants:
                if k8s.is_namespace_excluded(excluded_namespaces, robj):
                    continue      # (the engine reviews resultants through the same excluder stage as every object of the process)
                child = k8s.Review(obj=robj, ns=ns, source="Generated")          # :745-749
                for r in client.review(child, k8s.AUDIT_EP):
                    r = dict(r)
                    r["msg"] = (X.CHILD_MSG_PREFIX % tname) + " " + r["msg"]       # AggregateResponses
                    if action:
                        r["enforcementAction"] = action                           # OverrideEnforcementAction
                    results.append(r)
        for r in results:
            key = r["constraint"]
            totals[key] = totals.get(key, 0) + 1                        # :902
            by_action[r["enforcementAction"]] = by_action.get(r["enforcementAction"], 0) + 1
            g, ver, kind = k8s._gvk(obj)
            sv = {"group": g, "version": ver, "kind": kind, "namespace": k8s._meta(obj, "namespace"),
                  "name": k8s._meta(obj, "name"), "message": truncate_string(r["msg"]),
                  "enforcementAction": r["enforcementAction"],
                  "enforcementActions": r["scopedEnforcementActions"]}
            queues.setdefault(key, LimitQueue(limit)).push(sv)           # :925
            all_results.append({"object": idx, **r})
    return {"totals": totals, "by_action": by_action,
            "violations": {k: q.drain_descending() for k, q in queues.items()}, "results": all_results, "expand_errors": expand_errors}
