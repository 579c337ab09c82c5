#!/usr/bin/env python
"""Generates tests/golden/target_vectors.json from the reference's pkg/target tests by parsing the Go table literals and
the small helper calls they use (data only -- no reference code is copied):

  pkg/target/target_integration_test.go:164-413  TestConstraintEnforcement -- 26 scenarios through the REAL client + Rego driver
       with a deny-all template: `allowed` == no results, for three request shapes (Object, OldObject only,
       AugmentedUnstructured), :433-520
  pkg/controller/config/process/excluder_test.go:11-66  TestExactOrWildcardMatch -- the excluder's namespace patterns
  pkg/target/target_test.go:657-914              TestMatcher_Match -- Matcher.Match: object OR old object, review namespace
       vs cached namespace, error cases

Run in the authoring container only (needs /root/reference):  python tests/golden/make_target_vectors.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_match_vectors import CONST, P, REF, lex, val  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CONST.update({"types.SourceTypeDefault": "All"})


def gvk_of(n):
    d = {a[1]: val(b) for a, b in n[2]}
    return d.get("Group", ""), d.get("Version", ""), d.get("Kind", "")


def resource(n):
    """makeResource(gvk, name, labels...) / makeNamespacedResource(gvk, namespace, name, labels...) -- target_integration_test.go:135-152"""
    if n[0] == "ident" and n[1] == "nil":
        return None
    if n[0] == "call" and n[1] in ("matchedRawData", "unmatchedRawData", "namespacedRawData"):
        return RAW[n[1]](*[val(a) for a in n[2]])
    assert n[0] == "call" and n[1] in ("makeResource", "makeNamespacedResource"), n
    args = n[2]
    g, v, k = gvk_of(args[0])
    md = {}
    if n[1] == "makeNamespacedResource":
        md["namespace"] = val(args[1])
        md["name"] = val(args[2])
        rest = args[3:]
    else:
        md["name"] = val(args[1])
        rest = args[2:]
    if rest:
        md["labels"] = val(rest[0])
    # SetGroupVersionKind: apiVersion = GroupVersion.String() -> "group/version", or just "version" without a group
    api = (g + "/" + v) if g else v
    return {"apiVersion": api, "kind": k, "metadata": md}


def thing(group, kind, name, ns=None, labels=None):
    md = {"name": name}
    if ns:
        md["namespace"] = ns
    if labels:
        md["labels"] = labels
    return {"apiVersion": group + "/", "kind": kind, "metadata": md}


RAW = {   # target_test.go:635-655
    "matchedRawData": lambda: thing("some", "Thing", "bar", "foo", {"obj": "label"}),
    "unmatchedRawData": lambda: thing("another", "thing", "bar", "foo"),
    "namespacedRawData": lambda ns: thing("some", "Thing", "foo", ns, {"obj": "label"}),
}


def namespace(n):
    """makeNamespace(name, labels...) -- target_integration_test.go:154-162; &corev1.Namespace{ObjectMeta: ...}"""
    if n is None or (n[0] == "ident" and n[1] in ("nil",)):
        return None
    if n[0] == "ident" and n[1] == "ns":                      # `ns := makeNamespace("my-ns", {"ns": "label"})`, target_test.go:660
        return {"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": "my-ns", "labels": {"ns": "label"}}}
    if n[0] == "call":
        assert n[1] == "makeNamespace", n
        md = {"name": val(n[2][0])}
        if len(n[2]) > 1:
            md["labels"] = val(n[2][1])
        return {"apiVersion": "v1", "kind": "Namespace", "metadata": md}
    d = {a[1]: b for a, b in n[2]}
    meta = {a[1]: val(b) for a, b in d["ObjectMeta"][2]}
    md = {"name": meta.get("Name", "")}
    if "Labels" in meta:
        md["labels"] = meta["Labels"]
    return {"apiVersion": "v1", "kind": "Namespace", "metadata": md}


def constraint(n):
    """makeConstraint(set...(...)...) -- target_integration_test.go:46-133"""
    assert n[0] == "call" and n[1] == "makeConstraint", n
    m = {}
    for a in n[2]:
        fn, args = a[1], [val(x) for x in a[2]]
        if fn == "setKinds":
            m["kinds"] = [{"apiGroups": args[0], "kinds": args[1]}]
        elif fn == "setLabelSelector":
            m.setdefault("labelSelector", {}).setdefault("matchLabels", {})[args[0]] = args[1]
        elif fn == "setNamespaceSelector":
            m.setdefault("namespaceSelector", {}).setdefault("matchLabels", {})[args[0]] = args[1]
        elif fn == "setNamespaceName":
            m["namespaces"] = [args[0]]
        elif fn == "setExcludedNamespaceName":
            m["excludedNamespaces"] = [args[0]]
        elif fn == "setScope":
            m["scope"] = args[0]
        elif fn == "setSource":
            m["source"] = args[0]
        elif fn == "setName":
            m["name"] = args[0]
        else:
            raise SystemExit("unknown constraint option " + fn)
    c = {"apiVersion": "constraints.gatekeeper.sh/v1beta1", "kind": "DenyAll", "metadata": {"name": "my-constraint"}}
    if m:
        c["spec"] = {"match": m}
    return c


def balanced(src, j):
    """end index (exclusive) of the brace block that opens at src[j], skipping strings and comments"""
    depth, i = 0, j
    while i < len(src):
        c = src[i]
        if c == '"':
            i += 1
            while src[i] != '"':
                i += 2 if src[i] == "\\" else 1
        elif c == "`":
            i = src.index("`", i + 1)
        elif src.startswith("//", i):
            i = src.index("\n", i)
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise SystemExit("unbalanced braces")


def table(src, func_name, marker):
    i = src.index("func " + func_name + "(")
    j = src.index(marker, i)
    j = src.index("{", src.index("}", j))
    text = src[j:balanced(src, j)].replace("*makeResource(", "makeResource(")   # a dereferenced helper result is the same data
    return P(lex(text)).elems(), src[:j].count("\n") + 1


FOO_MATCH = {"source": "All", "kinds": [{"kinds": ["Thing"], "apiGroups": ["some"]}], "scope": "Namespaced", "namespaces": ["my-ns"],
             "labelSelector": {"matchLabels": {"obj": "label"}}, "namespaceSelector": {"matchLabels": {"ns": "label"}}}   # target_test.go:507-531
NSSEL_MATCH = {"namespaceSelector": {"matchLabels": {"ns": "label"}}}                                                      # target_test.go:497-505


def match_of(n):
    if n[0] == "ident" and n[1] == "nil":
        return None
    if n[0] == "call":
        return {"fooMatch": FOO_MATCH, "namespaceSelectorMatch": NSSEL_MATCH}[n[1]]
    from make_match_vectors import struct
    return struct(n, "match.Match")


def request(n):
    """The review shapes of TestMatcher_Match -> {object, oldObject, namespace (object), namespaceName, source, kind}"""
    if n[0] == "ident" and n[1] == "nil":
        return None
    if n[0] == "call":                                            # a bare *unstructured.Unstructured
        return {"shape": "Unstructured", "object": resource(n)}
    tp, fields = n[1], {a[1]: b for a, b in n[2]}
    out = {"shape": tp.split(".")[-1]}
    if out["shape"] == "AdmissionRequest":
        ar = fields
    elif out["shape"] == "AugmentedReview":
        out["namespace"] = namespace(fields.get("Namespace"))
        ar = {a[1]: b for a, b in fields["AdmissionRequest"][2]}
    elif out["shape"] == "AugmentedUnstructured":
        out["namespace"] = namespace(fields.get("Namespace"))
        o = fields["Object"]
        if o[0] == "lit":                                         # unstructured.Unstructured{Object: map{"key": "Some invalid json"}}
            out["object"] = {"key": "Some invalid json"}
        else:
            out["object"] = resource(o if o[0] == "call" else o)
        return out
    else:
        raise SystemExit("unknown request type " + tp)
    for key, name in (("Object", "object"), ("OldObject", "oldObject")):
        if key in ar:
            raw = {a[1]: b for a, b in ar[key][2]}["Raw"]
            if raw[0] == "ident" and raw[1] == "nsData":          # target_test.go:658: a v1 Namespace named foo
                out[name] = {"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": "foo"}}
            else:
                out[name] = resource(raw)
    if "Namespace" in ar:
        out["namespaceName"] = val(ar["Namespace"])
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    out = {"constraint_enforcement": [], "matcher_match": []}
    src = open(f"{REF}/pkg/target/target_integration_test.go").read()
    rows, _ = table(src, "TestConstraintEnforcement", "tcs := []struct")
    for _, row in rows:
        f = {a[1]: b for a, b in row[2]}
        out["constraint_enforcement"].append({
            "source_test": "pkg/target/target_integration_test.go:TestConstraintEnforcement", "name": val(f["name"]),
            "object": resource(f["obj"]), "namespace": namespace(f.get("ns")), "constraint": constraint(f["constraint"]),
            "allowed": val(f["allowed"])})
    src = open(f"{REF}/pkg/target/target_test.go").read()
    rows, _ = table(src, "TestMatcher_Match", "tests := []struct")
    for _, row in rows:
        f = {a[1]: b for a, b in row[2]}
        err = f.get("wantErr")
        out["matcher_match"].append({
            "source_test": "pkg/target/target_test.go:TestMatcher_Match", "name": val(f["name"]), "match": match_of(f["match"]),
            "cachedNamespace": namespace(f.get("cachedNs")), "request": request(f["req"]), "wantHandled": val(f["wantHandled"]),
            "wantErr": None if err is None or (err[0] == "ident" and err[1] == "nil") else err[1], "want": val(f["want"]) if "want" in f else False})
    src = open(f"{REF}/pkg/controller/config/process/excluder_test.go").read()
    rows, _ = table(src, "TestExactOrWildcardMatch", "tcs := []struct")
    out["excluder"] = []
    for _, row in rows:
        f = {a[1]: b for a, b in row[2]}
        out["excluder"].append({"source_test": "pkg/controller/config/process/excluder_test.go:TestExactOrWildcardMatch", "name": val(f["name"]),
                                "patterns": sorted(val(f["nsMap"]).keys()), "namespace": val(f["ns"]), "excluded": val(f["excluded"])})
    with open(os.path.join(HERE, "target_vectors.json"), "w") as fo:
        json.dump(out, fo, indent=1, sort_keys=True)
    print(len(out["constraint_enforcement"]), "enforcement scenarios,", len(out["matcher_match"]), "Matcher.Match vectors")


if __name__ == "__main__":
    main()
