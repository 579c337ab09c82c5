"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the in-tree Go parts of the
Review hot path: the `spec.match` pre-filter, wildcard globbing, the process excluder, enforcement-point
filtering, review construction and the `Client.Review` loop that ties them to the Rego evaluation.

Every function cites the reference file:line it restates (paths relative to /root/reference).  Pinned by the reference's
own vectors (tests/golden/match_vectors.json, wildcard_vectors.json, target_vectors.json -- extracted from
pkg/mutation/match/match_test.go, pkg/wildcard/wildcard_test.go, pkg/target/target_test.go, pkg/target/target_integration_test.go,
pkg/controller/config/process/excluder_test.go); the text of label-selector validation errors (apimachinery, un-vendored)
is parity unpinned.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import json
import re

from . import rego

TARGET_NAME = "admission.k8s.gatekeeper.sh"  # pkg/target/target.go:25

# enforcement points / actions -- pkg/util/enforcement_action.go:16-39
WEBHOOK_EP = "validation.gatekeeper.sh"
AUDIT_EP = "audit.gatekeeper.sh"
GATOR_EP = "gator.gatekeeper.sh"
VAP_EP = "vap.k8s.io"
ALL_EPS = "*"


class MatchError(Exception):
    """A matcher returned an error (tri-state: match / no match / error)."""


# ---------------------------------------------------------------------------------------------------
# pkg/wildcard/wildcard.go:17-41


def wildcard_matches(w: str, candidate: str) -> bool:
    """wildcard.Wildcard.Matches -- pkg/wildcard/wildcard.go:17-29"""
    if w.startswith("*") and w.endswith("*"):
        # strings.TrimSuffix(strings.TrimPrefix(w, "*"), "*"): for w == "*" the prefix trim leaves ""
        inner = w[1:]
        if inner.endswith("*"):
            inner = inner[:-1]
        return inner in candidate
    if w.startswith("*"):
        return candidate.endswith(w[1:])
    if w.endswith("*"):
        return candidate.startswith(w[:-1])
    return w == candidate


def wildcard_matches_generate_name(w: str, candidate: str) -> bool:
    """wildcard.Wildcard.MatchesGenerateName -- pkg/wildcard/wildcard.go:31-41"""
    if w.startswith("*") and w.endswith("*"):
        inner = w[1:]
        if inner.endswith("*"):
            inner = inner[:-1]
        return inner in candidate
    if w.endswith("*"):
        return candidate.startswith(w[:-1])
    return False


# ---------------------------------------------------------------------------------------------------
# k8s.io/apimachinery v0.35.4 (go.mod:43, NOT vendored) metav1.LabelSelectorAsSelector + labels.Selector,
# restated from its published behaviour; call sites pkg/mutation/match/match.go:87,110.

_NAME_RE = re.compile(r"^[A-Za-z0-9]([-A-Za-z0-9_.]*[A-Za-z0-9])?$")
_DNS1123_SUB = re.compile(r"^[a-z0-9]([-a-z0-9]*[a-z0-9])?(\.[a-z0-9]([-a-z0-9]*[a-z0-9])?)*$")


_QNAME_MSG = ("must consist of alphanumeric characters, '-', '_' or '.', and must start and end with an alphanumeric character "
              "(e.g. 'MyName',  or 'my.name',  or '123-abc', regex used for validation is '([A-Za-z0-9][-A-Za-z0-9_.]*)?[A-Za-z0-9]')")
_SUBDOMAIN_MSG = ("a lowercase RFC 1123 subdomain must consist of lower case alphanumeric characters, '-' or '.', and must start and end with "
                  "an alphanumeric character (e.g. 'example.com', regex used for validation is "
                  "'[a-z0-9]([-a-z0-9]*[a-z0-9])?(\\.[a-z0-9]([-a-z0-9]*[a-z0-9])?)*')")
_LABEL_VALUE_MSG = ("a valid label must be an empty string or consist of alphanumeric characters, '-', '_' or '.', and must start and end with "
                    "an alphanumeric character (e.g. 'MyValue',  or 'my_value',  or '12345', regex used for validation is "
                    "'(([A-Za-z0-9][-A-Za-z0-9_.]*)?[A-Za-z0-9])?')")


def _qualified_name_errors(k: str):
    """validation.IsQualifiedName (apimachinery util/validation): the messages, in order."""
    errs = []
    parts = k.split("/")
    if len(parts) == 1:
        name = parts[0]
    elif len(parts) == 2:
        prefix, name = parts
        if not prefix:
            errs.append("prefix part must be non-empty")
        else:
            if len(prefix) > 253:
                errs.append("prefix part must be no more than 253 characters")
            if not _DNS1123_SUB.match(prefix):
                errs.append("prefix part " + _SUBDOMAIN_MSG)
    else:
        return ["a qualified name " + _QNAME_MSG + " with an optional DNS subdomain prefix and '/' (e.g. 'example.com/MyName')"]
    if not name:
        errs.append("name part must be non-empty")
    elif len(name) > 63:
        errs.append("name part must be no more than 63 characters")
    if not _NAME_RE.match(name):
        errs.append("name part " + _QNAME_MSG)
    return errs


def _label_value_errors(v: str):
    """validation.IsValidLabelValue."""
    errs = []
    if len(v) > 63:
        errs.append("must be no more than 63 characters")
    if v != "" and not _NAME_RE.match(v):
        errs.append(_LABEL_VALUE_MSG)
    return errs


def _go_quote(x: str) -> str:
    return json.dumps(x, ensure_ascii=False)


def _go_strings(vals) -> str:
    """fmt %#v of a []string."""
    return "[]string(nil)" if not vals else "[]string{" + ", ".join(_go_quote(v) for v in vals) + "}"


def _requirement_error(key, op, vals):
    """labels.NewRequirement: every problem of one requirement as field errors, aggregated the way
    field.ErrorList.ToAggregate prints them (one: the message; several: [m1, m2])."""
    errs = []
    ke = _qualified_name_errors(key)
    if ke:
        errs.append("key: Invalid value: %s: %s" % (_go_quote(key), "; ".join(ke)))
    if op in ("In", "NotIn") and not vals:
        errs.append("values: Invalid value: %s: for 'in', 'notin' operators, values set can't be empty" % _go_strings(vals))
    if op in ("Exists", "DoesNotExist") and vals:
        errs.append("values: Invalid value: %s: values set must be empty for exists and does not exist" % _go_strings(vals))
    for i, v in enumerate(vals):
        ve = _label_value_errors(v)
        if ve:
            errs.append("values[%d][%s]: Invalid value: %s: %s" % (i, key, _go_quote(v), "; ".join(ve)))
    if not errs:
        return None
    uniq = list(dict.fromkeys(errs))
    return uniq[0] if len(uniq) == 1 else "[" + ", ".join(uniq) + "]"


def _json_brief(v):
    return json.dumps(int(v)) if isinstance(v, float) and v == int(v) else json.dumps(v)


def _go_type_name(v):
    return {bool: "bool", int: "number", float: "number", str: "string", list: "array", dict: "object", type(None): "null"}[type(v)]


def _selector_field_errors(sel):
    """convertToLabelSelector (a JSON round-trip into metav1.LabelSelector: type errors) followed by
    validation.ValidateLabelSelector(sel, {}, field.NewPath("spec", "labelSelector")) -- pkg/target/target.go:185-193,215-224
    and apimachinery's apis/meta/v1/validation.  Returns the list of messages; raises ValidateError for a type error."""
    P = "spec.labelSelector"
    ml = sel.get("matchLabels")
    if ml is not None and not isinstance(ml, dict):
        raise ValidateError("Could not convert JSON to LabelSelector: json: cannot unmarshal %s into Go struct field LabelSelector.matchLabels "
                            "of type map[string]string" % _go_type_name(ml))
    for k, v in (ml or {}).items():
        if not isinstance(v, str):
            raise ValidateError("Could not convert JSON to LabelSelector: json: cannot unmarshal %s into Go struct field LabelSelector.matchLabels "
                                "of type string" % _go_type_name(v))
    me = sel.get("matchExpressions")
    if me is not None and not isinstance(me, list):
        raise ValidateError("Could not convert JSON to LabelSelector: json: cannot unmarshal %s into Go struct field LabelSelector.matchExpressions "
                            "of type []v1.LabelSelectorRequirement" % _go_type_name(me))
    for x in me or []:
        if not isinstance(x, dict):
            raise ValidateError("Could not convert JSON to LabelSelector: json: cannot unmarshal %s into Go struct field LabelSelector.matchExpressions "
                                "of type v1.LabelSelectorRequirement" % _go_type_name(x))
        for f in ("key", "operator"):
            if x.get(f) is not None and not isinstance(x[f], str):
                raise ValidateError("Could not convert JSON to LabelSelector: json: cannot unmarshal %s into Go struct field LabelSelectorRequirement."
                                    "matchExpressions.%s of type string" % (_go_type_name(x[f]), f))
        vals = x.get("values")
        if vals is not None and (not isinstance(vals, list) or any(not isinstance(v, str) for v in vals)):
            raise ValidateError("Could not convert JSON to LabelSelector: json: cannot unmarshal %s into Go struct field LabelSelectorRequirement."
                                "matchExpressions.values of type []string" % _go_type_name(vals if not isinstance(vals, list) else
                                                                                             next(v for v in vals if not isinstance(v, str))))
    errs = []
    for k in sorted(ml or {}):
        for m in _qualified_name_errors(k):
            errs.append("%s.matchLabels: Invalid value: %s: %s" % (P, _go_quote(k), m))
        for m in _label_value_errors(ml[k]):
            errs.append("%s.matchLabels: Invalid value: %s: %s" % (P, _go_quote(ml[k]), m))
    for i, x in enumerate(me or []):
        fp = "%s.matchExpressions[%d]" % (P, i)
        op, vals, key = x.get("operator") or "", x.get("values") or [], x.get("key") or ""
        if op in ("In", "NotIn"):
            if not vals:
                errs.append(fp + ".values: Required value: must be specified when `operator` is 'In' or 'NotIn'")
        elif op in ("Exists", "DoesNotExist"):
            if vals:
                errs.append(fp + ".values: Forbidden: may not be specified when `operator` is 'Exists' or 'DoesNotExist'")
        else:
            errs.append(fp + ".operator: Invalid value: %s: not a valid selector operator" % _go_quote(op))
        for m in _qualified_name_errors(key):
            errs.append(fp + ".key: Invalid value: %s: %s" % (_go_quote(key), m))
        for j, v in enumerate(vals):
            for m in _label_value_errors(v):
                errs.append(fp + ".values[%d]: Invalid value: %s: %s" % (j, _go_quote(v), m))
    return errs


class ValidateError(Exception):
    pass


def to_matcher_error(constraint):
    """K8sValidationTarget.ToMatcher / convertToMatch (pkg/target/target.go:226-254): spec.match must be a map whose members have
    the JSON types of match.Match (pkg/mutation/match/match_types.go:13-51).  Returns the error text ("unable to create
    matcher: ...", ErrCreatingMatcher) or None; pinned by TestToMatcher (pkg/target/target_test.go:544-626)."""
    spec = constraint.get("spec") if isinstance(constraint, dict) else None
    if not isinstance(spec, dict) or spec.get("match") is None:
        return None
    m = spec["match"]
    pre = "unable to create matcher: "
    if not isinstance(m, dict):
        return pre + ".spec.match accessor error: %s is of the type %s, expected map[string]interface{}" % (_json_brief(m), _go_type_name(m))

    def bad(v, field, typ):
        return pre + "could not convert JSON to Match: json: cannot unmarshal %s into Go struct field Match.%s of type %s" % (_go_type_name(v), field, typ)

    def strlist(v, field, typ):
        if v is None:
            return None
        if not isinstance(v, list):
            return bad(v, field, typ)
        for x in v:
            if not isinstance(x, str):
                return bad(x, field, "string")
        return None
    for f in ("source", "scope", "name"):
        if m.get(f) is not None and not isinstance(m[f], str):
            return bad(m[f], f, "string")
    kinds = m.get("kinds")
    if kinds is not None:
        if not isinstance(kinds, list):
            return bad(kinds, "kinds", "[]match.Kinds")
        for k in kinds:
            if not isinstance(k, dict):
                return bad(k, "kinds", "match.Kinds")
            for f in ("apiGroups", "kinds"):
                e = strlist(k.get(f), "kinds." + f, "[]string")
                if e:
                    return e
    for f in ("namespaces", "excludedNamespaces"):
        e = strlist(m.get(f), f, "[]wildcard.Wildcard")
        if e:
            return e
    for f in ("labelSelector", "namespaceSelector"):
        sel = m.get(f)
        if sel is None:
            continue
        if not isinstance(sel, dict):
            return bad(sel, f, "v1.LabelSelector")
        try:
            _selector_field_errors(sel)
        except ValidateError as e:
            return pre + str(e).replace("Could not convert JSON to LabelSelector", "could not convert JSON to Match")
    return None


def validate_constraint(constraint):
    """K8sValidationTarget.ValidateConstraint (pkg/target/target.go:178-214): the frameworks client calls it before
    Driver.AddConstraint.  Raises ValidateError; pinned by TestValidateConstraint (pkg/target/target_test.go:42-399)."""
    cur = constraint
    for part in ("spec", "match"):
        if not isinstance(cur, dict) or part not in cur:
            return
        cur = cur[part]
    if not isinstance(cur, dict):
        raise ValidateError(".spec.match accessor error: %s is of the type %s, expected map[string]interface{}" % (_json_brief(cur), _go_type_name(cur)))
    for f in ("labelSelector", "namespaceSelector"):
        if f not in cur or cur[f] is None:
            continue
        sel = cur[f]
        if not isinstance(sel, dict):
            raise ValidateError(".spec.match.%s accessor error: %s is of the type %s, expected map[string]interface{}" % (f, _json_brief(sel), _go_type_name(sel)))
        errs = _selector_field_errors(sel)
        if errs:
            uniq = list(dict.fromkeys(errs))
            raise ValidateError(uniq[0] if len(uniq) == 1 else "[" + ", ".join(uniq) + "]")


def label_selector_requirements(sel):
    """LabelSelectorAsSelector: returns a list of (key, op, values) or raises MatchError.  Requirements are built in
    order (matchLabels -- sorted here, a Go map there -- then matchExpressions) and the first failing one is the error."""
    reqs = []
    if sel is None:
        return None
    for k, v in sorted((sel.get("matchLabels") or {}).items()):
        e = _requirement_error(k, "In", [v])
        if e:
            raise MatchError(e)
        reqs.append((k, "In", [v]))
    for x in sel.get("matchExpressions") or []:
        op = x.get("operator", "")
        vals = list(x.get("values") or [])
        if op not in ("In", "NotIn", "Exists", "DoesNotExist"):
            raise MatchError(f'"{op}" is not a valid label selector operator')
        e = _requirement_error(x.get("key", ""), op, vals)
        if e:
            raise MatchError(e)
        reqs.append((x.get("key", ""), op, vals))
    return reqs


def selector_matches(reqs, labels) -> bool:
    """labels.Selector.Matches: all requirements ANDed; NotIn is true when the key is absent."""
    labels = labels or {}
    for k, op, vals in reqs:
        has = k in labels
        if op == "In":
            if not (has and labels[k] in vals):
                return False
        elif op == "NotIn":
            if has and labels[k] in vals:
                return False
        elif op == "Exists":
            if not has:
                return False
        elif op == "DoesNotExist":
            if has:
                return False
    return True


# ---------------------------------------------------------------------------------------------------
# object accessors (unstructured.Unstructured getters)


def _gvk(obj):
    api = obj.get("apiVersion", "") or ""
    if not isinstance(api, str):
        api = ""
    if "/" in api:
        g, v = api.split("/", 1)
    else:
        g, v = "", api
    kind = obj.get("kind", "")
    return g, v, kind if isinstance(kind, str) else ""


def _meta(obj, field):
    md = obj.get("metadata")
    if not isinstance(md, dict):
        return ""
    v = md.get(field, "")
    return v if isinstance(v, str) else ""


def _labels(obj):
    md = obj.get("metadata")
    if not isinstance(md, dict):
        return {}
    ls = md.get("labels")
    if not isinstance(ls, dict):
        return {}
    # unstructured.GetLabels -> NestedStringMap: any non-string value makes the whole map unreadable
    if any(not isinstance(v, str) for v in ls.values()):
        return {}
    return ls


def is_namespace(obj) -> bool:
    """match.IsNamespace -- pkg/mutation/match/match.go:255-258"""
    g, _, k = _gvk(obj)
    return k == "Namespace" and g == ""


# ---------------------------------------------------------------------------------------------------
# pkg/mutation/match/match.go:32-268


def _ns_name_for(match_list_empty, obj, ns):
    """Shared name selection of namespacesMatch/excludedNamespacesMatch -- match.go:118-179.
    Returns (decided, value): decided=True means `value` is the final answer's "no name" case."""
    if is_namespace(obj):
        return _meta(obj, "name")
    if ns is not None:
        return _meta(ns, "name")
    if _meta(obj, "namespace") != "":
        return _meta(obj, "namespace")
    return None


def matches(match, obj, ns, source) -> bool:
    """match.Matches -- pkg/mutation/match/match.go:32-65: AND of 8 matchers in fixed order with early
    exit; raises MatchError where the Go code returns an error."""
    if obj is None:
        raise MatchError("failed to run Match criteria: obj must be non-nil")  # :33-38
    try:
        for fn in (_kinds, _scope, _namespaces, _excluded_namespaces, _label_selector, _namespace_selector,
                   _names, _source):
            if not fn(match, obj, ns, source):
                return False
    except MatchError as e:
        raise MatchError(f"failed to run Match criteria: {e}")  # :52-54
    return True


def _kinds(m, obj, ns, src):  # match.go:181-201
    kinds = m.get("kinds") or []
    if not kinds:
        return True
    g, _, k = _gvk(obj)
    for kk in kinds:
        ks = kk.get("kinds") or []
        gs = kk.get("apiGroups") or []
        if not (len(ks) == 0 or "*" in ks or k in ks):
            continue
        if len(gs) == 0 or "*" in gs or g in gs:
            return True
    return False


def _scope(m, obj, ns, src):  # match.go:214-227
    has_ns = _meta(obj, "namespace") != "" or ns is not None
    is_ns = is_namespace(obj)
    scope = m.get("scope", "")
    if scope == "Cluster":
        return is_ns or not has_ns
    if scope == "Namespaced":
        return (not is_ns) and has_ns
    return True


def _namespaces(m, obj, ns, src):  # match.go:150-179
    pats = m.get("namespaces") or []
    if not pats:
        return True
    name = _ns_name_for(False, obj, ns)
    if name is None:
        return True
    return any(wildcard_matches(p, name) for p in pats)


def _excluded_namespaces(m, obj, ns, src):  # match.go:118-148
    pats = m.get("excludedNamespaces") or []
    if not pats:
        return True
    name = _ns_name_for(False, obj, ns)
    if name is None:
        return True
    return not any(wildcard_matches(p, name) for p in pats)


def _label_selector(m, obj, ns, src):  # match.go:103-116
    sel = m.get("labelSelector")
    if sel is None:
        return True
    reqs = label_selector_requirements(sel)
    return selector_matches(reqs, _labels(obj))


def _namespace_selector(m, obj, ns, src):  # match.go:73-101
    sel = m.get("namespaceSelector")
    if sel is None:
        return True
    is_ns = is_namespace(obj)
    if not is_ns and ns is None and _meta(obj, "namespace") == "":
        return True
    reqs = label_selector_requirements(sel)
    if is_ns:
        return selector_matches(reqs, _labels(obj))
    if ns is None:
        raise MatchError("namespace selector for namespace-scoped object but missing Namespace")
    return selector_matches(reqs, _labels(ns))


def _names(m, obj, ns, src):  # match.go:203-212
    name = m.get("name", "") or ""
    if name == "":
        return True
    return wildcard_matches(name, _meta(obj, "name")) or wildcard_matches_generate_name(
        name, _meta(obj, "generateName"))


_VALID_SOURCES = ("All", "Generated", "Original")  # pkg/mutation/types/mutator.go:14-27


def _source(m, obj, ns, src):  # match.go:229-253
    msrc = m.get("source", "") or ""
    if msrc == "":
        msrc = "All"
    elif msrc not in _VALID_SOURCES:
        raise MatchError(f'invalid source field "{msrc}"')
    if (src or "") == "" and msrc != "All":
        raise MatchError(f"source field not specified for resource {_meta(obj, 'name')}")
    if msrc == "All":
        return True
    if src not in _VALID_SOURCES:
        raise MatchError(f'invalid source field "{src}"')
    return msrc == src


# ---------------------------------------------------------------------------------------------------
# pkg/target/matcher.go:21-93  (Matcher.Match / matchAny)


class Review:
    """gkReview (pkg/target/review.go:16-29) after HandleReview normalisation."""

    __slots__ = ("obj", "old", "ns", "source", "kind", "name", "namespace", "operation", "user_info",
                 "is_admission")

    def __init__(self, obj=None, old=None, ns=None, source="", operation="", user_info=None, kind=None,
                 name=None, namespace=None, is_admission=False):
        self.obj, self.old, self.ns, self.source = obj, old, ns, source
        self.operation, self.user_info, self.is_admission = operation, user_info or {}, is_admission
        ref = obj if obj is not None else old
        if kind is None and ref is not None:
            g, v, k = _gvk(ref)
            kind = {"group": g, "version": v, "kind": k}
        self.kind = kind or {"group": "", "version": "", "kind": ""}
        self.name = name if name is not None else (_meta(ref, "name") if ref else "")
        self.namespace = namespace if namespace is not None else (_meta(ref, "namespace") if ref else "")


def handle_review(review: Review) -> Review:
    """K8sValidationTarget.handleReview tail -- setObjectOnDelete, pkg/target/target.go:262-280."""
    if review.operation == "DELETE":
        if review.old is None:
            raise ValueError("oldObject cannot be nil for DELETE operations")
        review.obj = review.old
    return review


def matcher_match(match, review: Review, ns_cache=None) -> bool:
    """Matcher.Match + matchAny -- pkg/target/matcher.go:21-71."""
    if match is None:
        return True  # :22-25 no-op if Match unspecified
    ns = review.ns
    if ns is None and review.namespace != "" and ns_cache is not None:
        ns = ns_cache.get(review.namespace)  # :37-39
    nil = 0
    for o in (review.obj, review.old):
        if o is None:
            nil += 1
            continue
        try:
            if matches(match, o, ns, review.source):
                return True
        except MatchError as e:
            # fmt.Errorf("%w: %v :%w", ErrMatching, obj.GetName(), err) -- matcher.go:58-60
            raise MatchError(f"error matching the requested object: {_meta(o, 'name')} :{e}")
    if nil == 2:
        raise MatchError("invalid request object: neither object nor old object are defined")
    return False


# ---------------------------------------------------------------------------------------------------
# pkg/controller/config/process/excluder.go:95-127


def is_namespace_excluded(excluded_patterns, obj) -> bool:
    name = _meta(obj, "name") if is_namespace(obj) else _meta(obj, "namespace")
    return any(wildcard_matches(p, name) for p in excluded_patterns)


# ---------------------------------------------------------------------------------------------------
# enforcement actions -- pkg/util/enforcement_action.go:132-174


def get_enforcement_action(constraint) -> str:
    ea = (constraint.get("spec") or {}).get("enforcementAction", "") or ""
    if ea == "":
        return "deny"
    return ea if ea in ("deny", "dryrun", "warn", "scoped") else "unrecognized"


def scoped_actions_for_ep(ep: str, constraint):
    out = []
    for sea in (constraint.get("spec") or {}).get("scopedEnforcementActions") or []:
        for p in sea.get("enforcementPoints") or []:
            if p.get("name") == ep or p.get("name") == ALL_EPS:
                out.append(sea.get("action", ""))
                break
    return out


# ---------------------------------------------------------------------------------------------------
# input.review document -- shape of gkReview as JSON (k8s.io/api admission/v1
# AdmissionRequest JSON tags, module not vendored).  `namespaceObject` is injected from
# reviews.Namespace(nsMap) (pkg/util/namespace.go:15-25, website/docs/input.md:6-16).


def review_document(review: Review):
    doc = {
        "uid": "",
        "kind": dict(review.kind),
        "resource": {"group": "", "version": "", "resource": ""},
        "operation": review.operation or "",
        "userInfo": dict(review.user_info or {}),
        "object": review.obj,
        "oldObject": review.old if review.operation != "DELETE" or review.old is not review.obj else review.old,
        "options": None,
    }
    if review.name:
        doc["name"] = review.name
    if review.namespace:
        doc["namespace"] = review.namespace
    if review.ns is not None:
        doc["namespaceObject"] = review.ns
    return doc


# ---------------------------------------------------------------------------------------------------
# Client (frameworks/constraint pkg/client, NOT vendored): AddTemplate / AddConstraint / Review restated
# from its call sites and the reference tests that pin its behaviour.


class Template:
    def __init__(self, kind, rego_src, libs=()):
        self.kind = kind
        self.module = rego.Module(rego_src, libs)


def template_from_yaml_obj(ct):
    """ConstraintTemplate dict -> (kind, rego source) or, for a template with libs, (kind, rego source, libs).  Legacy
    `targets[].rego` / `targets[].libs` are surfaced as engine "Rego" (pkg/fakes/fixtures.go:32-44); a `code` entry for the
    Rego engine wins over them."""
    kind = ct["spec"]["crd"]["spec"]["names"]["kind"]
    tgt = ct["spec"]["targets"][0]
    src = libs = None
    for c in tgt.get("code") or []:
        if c.get("engine") == "Rego":
            src = (c.get("source") or {}).get("rego")
            libs = (c.get("source") or {}).get("libs")
    if not src:
        src, libs = tgt.get("rego"), tgt.get("libs")
    if not src:
        raise rego.RegoError("no Rego source for template (ErrNoDriver)")
    return (kind, src, tuple(libs)) if libs else (kind, src)


class Client:
    """Sequential, one-object-at-a-time evaluation -- exactly the shape of the reference's audit loop
    (pkg/audit/manager.go:686-720) around Client.Review."""

    def __init__(self):
        self.templates = {}     # kind -> Template
        self.constraints = {}   # (kind, name) -> constraint dict ; insertion-ordered
        self.ns_cache = {}      # name -> namespace object   (pkg/target/ns_cache.go:15-85)
        self.inventory = {}     # data.inventory: {"cluster": {gv: {kind: {name: obj}}}, "namespace": {ns: {gv: {kind: {name: obj}}}}}
        self._inventory_doc = None

    def add_template(self, kind, rego_src, libs=()):
        self.templates[kind] = Template(kind, rego_src, libs)

    def add_constraint(self, constraint):
        kind = constraint["kind"]
        if kind not in self.templates:
            raise KeyError(f"no template for constraint kind {kind}")
        terr = to_matcher_error(constraint)
        if terr:
            raise MatchError(terr)
        m = (constraint.get("spec") or {}).get("match")
        if m is not None:
            # ValidateConstraint (pkg/target/target.go:178-214) rejects bad selectors at load time
            for f in ("labelSelector", "namespaceSelector"):
                if m.get(f) is not None:
                    try:
                        label_selector_requirements(m[f])
                    except MatchError:
                        pass  # invalid operators are only caught at match time (match_test.go:388-405)
        self.constraints[(kind, constraint["metadata"]["name"])] = constraint

    def remove_constraint(self, kind, name):
        self.constraints.pop((kind, name), None)

    def add_namespace(self, ns_obj, name=None):
        """nsCache.Add (pkg/target/ns_cache.go:22-44; TestNamespaceCache, pkg/target/target_test.go:983-1153): a non-map is an
        error, a map that is not a core Namespace is ignored, a Namespace that does not convert to corev1.Namespace is an error."""
        if not isinstance(ns_obj, dict):
            raise MatchError("cannot cache non-namespace type: cannot cache type %s, want map[string]interface {}" % _go_type_name(ns_obj))
        g, _, k = _gvk(ns_obj)
        if k != "Namespace" or g:
            return
        md = ns_obj.get("metadata")
        ok = all(x is None or isinstance(x, dict) for x in (md, ns_obj.get("spec"), ns_obj.get("status")))
        if ok and isinstance(md, dict):
            for f in ("labels", "annotations"):
                ls = md.get(f)
                ok = ok and (ls is None or (isinstance(ls, dict) and all(isinstance(v, str) for v in ls.values())))
            ok = ok and (md.get("name") is None or isinstance(md.get("name"), str))
        if not ok:
            raise MatchError("cannot cache non-namespace type: cannot cache Namespace: <nil>")
        self.ns_cache[name if name is not None else _meta(ns_obj, "name")] = ns_obj

    @staticmethod
    def data_path(obj):
        """processUnstructured (pkg/target/target.go:40-57): where a synced object lives under data.inventory."""
        api = obj.get("apiVersion") or ""
        group, _, version = api.rpartition("/")
        name = _meta(obj, "name") or ""
        if not version:
            raise ValueError("invalid request object: resource %s has no version" % name)
        if not obj.get("kind"):
            raise ValueError("invalid request object: resource %s has no kind" % name)
        gv = group + "/" + version if group else version
        ns = _meta(obj, "namespace") or ""
        return ["cluster", gv, obj["kind"], name] if not ns else ["namespace", ns, gv, obj["kind"], name]

    def add_data(self, obj, path=None):
        """Client.AddData: the object is stored at its path for referential templates; Namespaces also enter the nsCache
        (the caller does that through add_namespace, as the reference's Client does through cache.Add)."""
        path = list(path) if path else self.data_path(obj)
        cur = self.inventory
        for p in path[:-1]:
            cur = cur.setdefault(p, {})
        cur[path[-1]] = obj
        self._inventory_doc = None

    def remove_data(self, path):
        cur = self.inventory
        for p in path[:-1]:
            cur = cur.get(p)
            if cur is None:
                return
        cur.pop(path[-1], None)
        self._inventory_doc = None

    def data_doc(self):
        if self._inventory_doc is None:
            self._inventory_doc = rego.from_json({"inventory": self.inventory})
        return self._inventory_doc

    def review(self, review: Review, enforcement_point: str = AUDIT_EP):
        """Returns a list of result dicts {constraint:(kind,name), msg, details, enforcementAction,
        scopedEnforcementActions}.  Matcher errors become results carrying the error text (pinned by
        pkg/gator/verify/runner_test.go:986-989: `Violations: yes, Message: "missing Namespace"`)."""
        review = handle_review(review)
        doc = rego.from_json(review_document(review))
        results = []
        for (kind, name), c in self.constraints.items():
            action = get_enforcement_action(c)
            scoped = None
            if action == "scoped":
                scoped = scoped_actions_for_ep(enforcement_point, c)
                if not scoped:
                    continue  # constraint not enforced at this enforcement point
            spec = c.get("spec") or {}
            try:
                if not matcher_match(spec.get("match"), review, self.ns_cache):
                    continue
            except MatchError as e:
                # the frameworks client words the autoreject result "unable to match constraints: <matcher error>" -- pinned by
                # test/gator/test/test.bats:276
                results.append({"constraint": (kind, name), "msg": "unable to match constraints: " + str(e), "details": {},
                                "enforcementAction": action, "scopedEnforcementActions": scoped or [],
                                "autoreject": True})
                continue
            inp = rego.RObj({"review": doc, "parameters": rego.from_json(spec.get("parameters") or {})})
            for v in rego.eval_violations(self.templates[kind].module, inp, self.data_doc()):
                results.append({"constraint": (kind, name), "msg": v["msg"],
                                "details": rego.to_json(v["details"]) if "details" in v else None,
                                "enforcementAction": action, "scopedEnforcementActions": scoped or []})
        return results


# ---------------------------------------------------------------------------------------------------
# admission messages -- pkg/webhook/policy.go:238-355 (getValidationMessages)

SUPPORTED_ACTIONS = ("deny", "dryrun", "warn")     # pkg/util/enforcement_action.go:60-70


def validation_messages(results):
    """results: the Client.Review results of ONE admission request at the webhook enforcement point.
    Returns (denyMsgs, warnMsgs): scoped results use their actions for the enforcement point (unsupported ones
    skipped; none left => the result is dropped), others their own action (unsupported => dropped)."""
    deny, warn = [], []
    for r in results:
        if r["enforcementAction"] == "scoped":
            actions = [a for a in r["scopedEnforcementActions"] if a in SUPPORTED_ACTIONS]
            if not actions:
                continue
        else:
            if r["enforcementAction"] not in SUPPORTED_ACTIONS:
                continue
            actions = [r["enforcementAction"]]
        name = r["constraint"][1] if isinstance(r["constraint"], tuple) else r["constraint"].split("/", 1)[1]
        for a in actions:
            if a == "deny":
                deny.append("[%s] %s" % (name, r["msg"]))
            if a == "warn":
                warn.append("[%s] %s" % (name, r["msg"]))
    return deny, warn
