"""TEST INFRASTRUCTURE: CPU restatement of the reference's expansion of generator resources (pkg/expansion/system.go,
aggregate.go) -- ExpansionTemplates turn a workload object (a Deployment) into the resources it implies (a Pod) so that
constraints written for the implied kind are evaluated too; results are reported on the parent with an "[Implied by <template>]"
prefix and an optional enforcement-action override.  Mutators (pkg/mutation) are outside this engine's scope: resultants are not
mutated.  Pinned by pkg/expansion/system_test.go (TestExpand / TestExpandResource entries without mutators) and
test/gator/test/test.bats:268-289 via tests/golden/expansion_vectors.json."""
import copy

MAX_RECURSION_DEPTH = 30          # system.go:30
CHILD_MSG_PREFIX = "[Implied by %s]"   # aggregate.go:11


class ExpansionError(Exception):
    pass


def _gvk(obj):
    av = obj.get("apiVersion") or ""
    if "/" in av:
        g, v = av.split("/", 1)
    else:
        g, v = "", av
    return g, v, obj.get("kind") or ""


def validate_template(t):
    """ValidateTemplate -- system.go:85-112"""
    name = (t.get("metadata") or {}).get("name") or ""
    spec = t.get("spec") or {}
    if not name:
        raise ExpansionError("ExpansionTemplate has empty name field")
    if len(name) >= 64:
        raise ExpansionError("ExpansionTemplate name must be less than 64 characters")
    if not spec.get("templateSource"):
        raise ExpansionError("ExpansionTemplate %s has empty source field" % name)
    gen = spec.get("generatedGVK") or {}
    if not (gen.get("group") or gen.get("version") or gen.get("kind")):
        raise ExpansionError("ExpansionTemplate %s has empty generatedGVK field" % name)
    if not spec.get("applyTo"):
        raise ExpansionError("ExpansionTemplate %s must specify ApplyTo" % name)
    g = (gen.get("group") or "", gen.get("version") or "", gen.get("kind") or "")
    for a in spec["applyTo"]:
        if _apply_matches(a, g):
            raise ExpansionError("ExpansionTemplate %s generates GVK %s/%s, Kind=%s, but also applies to that same GVK" % ((name,) + g))   # %v of a schema.GroupVersionKind


def _apply_matches(a, gvk):
    """ApplyTo.Matches -- pkg/mutation/match/apply_to.go:45-57"""
    return gvk[0] in (a.get("groups") or []) and gvk[1] in (a.get("versions") or []) and gvk[2] in (a.get("kinds") or [])


def _name(obj):
    """Unstructured.GetName(): metadata.name when it is a string, else \"\" """
    md = obj.get("metadata") if isinstance(obj, dict) else None
    n = md.get("name") if isinstance(md, dict) else None
    return n if isinstance(n, str) else ""


def expand_resource(obj, ns_name, template):
    """expandResource -- system.go:203-247.  `ns_name`: name of the review's Namespace object, or None."""
    spec = template.get("spec") or {}
    src_path = spec.get("templateSource") or ""
    if not src_path:
        raise ExpansionError("cannot expand resource using a template with no source")
    gen = spec.get("generatedGVK") or {}
    if not (gen.get("group") or gen.get("version") or gen.get("kind")):
        raise ExpansionError("cannot expand resource using template with empty generatedGVK")
    # unstructured.NestedMap (system.go:225-231): a missing key is "not found"; walking INTO something that is not a map, or a
    # value at the end that is not a map, is an accessor error
    cur = obj
    for key in src_path.split("."):
        if not isinstance(cur, dict):
            raise ExpansionError("could not extract source field from unstructured")
        if key not in cur:
            raise ExpansionError('could not find source field "%s" in resource %s' % (src_path, _name(obj)))
        cur = cur[key]
    if not isinstance(cur, dict):
        raise ExpansionError("could not extract source field from unstructured")
    res = copy.deepcopy(cur)
    group, version, kind = gen.get("group") or "", gen.get("version") or "", gen.get("kind") or ""
    res["apiVersion"] = (group + "/" + version) if group else version
    res["kind"] = kind
    md = res.setdefault("metadata", {})
    if not isinstance(md, dict):
        md = res["metadata"] = {}
    if ns_name is not None:
        md["namespace"] = ns_name
    else:
        # unstructured.NestedString(obj, "metadata", "namespace") (system.go:239-246): absent is fine (a cluster-scoped parent),
        # present but not a string -- or a metadata that is not a map -- is an error
        pmd = obj.get("metadata") if isinstance(obj, dict) else None
        if pmd is not None and not isinstance(pmd, dict):
            raise ExpansionError('could not extract namespace field "%s" in parent resource %s' % (src_path, _name(obj)))
        if isinstance(pmd, dict) and "namespace" in pmd:
            if not isinstance(pmd["namespace"], str):
                raise ExpansionError('could not extract namespace field "%s" in parent resource %s' % (src_path, _name(obj)))
            md["namespace"] = pmd["namespace"]
    pname = (obj.get("metadata") or {}).get("name") or ""
    md["name"] = (pname + ("-" if kind else "") + kind).lower()          # mockNameForResource -- system.go:289-297
    # ensureOwnerReference -- system.go:251-283
    pav, pk = obj.get("apiVersion") or "", obj.get("kind") or ""
    if pav and pk and pname:
        refs = md.get("ownerReferences") or []
        if not any(isinstance(r, dict) and r.get("apiVersion") == pav and r.get("kind") == pk and r.get("name") == pname for r in refs):
            md["ownerReferences"] = list(refs) + [{"apiVersion": pav, "kind": pk, "name": pname, "uid": ""}]
    return res


class System:
    def __init__(self):
        self.templates = {}

    def upsert(self, t):
        """System.UpsertTemplate -> db.upsert (system.go:56-66, db.go:222-245): a template that closes a cycle is stored all the same
        (set aside with the others on the cycle) and reported"""
        validate_template(t)
        self.templates[t["metadata"]["name"]] = t
        if t["metadata"]["name"] in self._conflicted():
            raise ExpansionError("template forms expansion cycle")

    def conflicts(self):
        """System.GetConflicts -- system.go:81-83"""
        return self._conflicted()

    def remove(self, name):
        self.templates.pop(name, None)

    def _conflicted(self):
        """templates on a cycle of the generatedGVK -> applyTo graph are set aside (db.go: hasConflicts)"""
        names = sorted(self.templates)
        edges = {n: [] for n in names}
        for a in names:
            ga = self.templates[a]["spec"]["generatedGVK"]
            g = (ga.get("group") or "", ga.get("version") or "", ga.get("kind") or "")
            for b in names:
                if any(_apply_matches(x, g) for x in self.templates[b]["spec"]["applyTo"]):
                    edges[a].append(b)
        bad = set()
        for s in names:
            stack, seen = [(s, iter(edges[s]))], {s}
            while stack:
                node, it = stack[-1]
                nxt = next(it, None)
                if nxt is None:
                    stack.pop()
                    continue
                if nxt == s:
                    bad.add(s)
                    break
                if nxt not in seen:
                    seen.add(nxt)
                    stack.append((nxt, iter(edges[nxt])))
        return bad

    def templates_for(self, gvk):
        bad = self._conflicted()
        return [self.templates[n] for n in sorted(self.templates)
                if n not in bad and any(_apply_matches(a, gvk) for a in self.templates[n]["spec"]["applyTo"])]

    def expand(self, obj, ns_name=None, depth=0):
        """System.Expand / expandRecursive -- system.go:137-167: [(resultant, template name, enforcementAction)], grandchildren first"""
        if depth >= MAX_RECURSION_DEPTH:
            raise ExpansionError("maximum recursion depth of %d reached" % MAX_RECURSION_DEPTH)
        gvk = _gvk(obj)
        if gvk == ("", "", ""):
            raise ExpansionError("cannot expand resource %s with empty GVK" % (obj.get("metadata") or {}).get("name", ""))
        res = [(expand_resource(obj, ns_name, t), t["metadata"]["name"], (t.get("spec") or {}).get("enforcementAction") or "")
               for t in self.templates_for(gvk)]
        out = []
        for r in res:
            out.extend(self.expand(r[0], ns_name, depth + 1))
        return out + res


def review_with_expansion(client, system, review, enforcement_point):
    """What the audit loop / the webhook do around Client.Review (pkg/audit/manager.go:733-765, pkg/webhook/policy.go:610-646):
    review the object, expand it, review every resultant as a Generated resource with the same Namespace, override the
    enforcement action where the template says so, prefix the child messages, append them to the parent's results."""
    from oracle import k8s
    results = list(client.review(review, enforcement_point))
    # the generator is the request's object -- the OLD object of a DELETE (getReqObject, pkg/webhook/policy.go:435-440,599-603)
    obj = review.old if review.operation == "DELETE" else review.obj
    if not isinstance(obj, dict):
        return results
    if review.namespace is not None:      # an admission request: obj.SetNamespace(req.Namespace) before expanding (policy.go:608)
        obj = dict(obj)
        md = dict(obj.get("metadata") or {}) if isinstance(obj.get("metadata"), dict) or obj.get("metadata") is None else {}
        md.pop("namespace", None)
        if review.namespace:
            md["namespace"] = review.namespace
        obj["metadata"] = md
    ns_name = None
    if review.ns is not None:
        ns_name = (review.ns.get("metadata") or {}).get("name", "")
    else:
        key = review.namespace or (obj.get("metadata") or {}).get("namespace") or ""
        if key and key in client.ns_cache:
            ns_name = (client.ns_cache[key].get("metadata") or {}).get("name", "")
    for robj, tname, action in system.expand(obj, ns_name):
        child = k8s.Review(obj=robj, ns=review.ns, source="Generated", namespace=review.namespace if review.namespace is not None else None)
        for r in client.review(child, enforcement_point):
            r = dict(r)
            r["msg"] = (CHILD_MSG_PREFIX % tname) + " " + r["msg"]
            if action:
                r["enforcementAction"] = action      # OverrideEnforcementAction (aggregate.go:47-58): only this field; the scoped list stays
            results.append(r)
    return results
