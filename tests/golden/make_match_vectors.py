#!/usr/bin/env python
"""Generates tests/golden/match_vectors.json and tests/golden/wildcard_vectors.json from the reference's
own table-driven Go tests, by parsing the Go composite literals (data only -- no reference code is copied):

  pkg/mutation/match/match_test.go:17-683   TestMatch        (Matches vectors incl. error cases)
  pkg/mutation/match/match_test.go:847-1040 Test_namesMatch  (name / generateName vectors)
  pkg/wildcard/wildcard_test.go:7-193       TestMatches / TestMatchesGenerateName

Run in the authoring container only (needs /root/reference):  python tests/golden/make_match_vectors.py
"""
import json
import os
import re
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

TOK = re.compile(r'\s+|//[^\n]*|(?P<str>"(?:[^"\\]|\\.)*")|(?P<raw>`[^`]*`)|(?P<num>\d+)|(?P<id>[A-Za-z_][A-Za-z0-9_]*)'
                 r'|(?P<op>\.\.\.|:=|[-&*{}()\[\],.:=!<>|+])')


def lex(src):
    out = []
    i = 0
    while i < len(src):
        m = TOK.match(src, i)
        if not m:
            raise SystemExit(f"lex error at {src[i:i+30]!r}")
        i = m.end()
        if m.lastgroup:
            out.append((m.lastgroup, m.group()))
    return out


class P:
    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        x = self.t[self.i]
        self.i += 1
        return x

    def accept(self, v):
        if self.peek()[1] == v:
            self.i += 1
            return True
        return False

    def expect(self, v):
        x = self.next()
        assert x[1] == v, (x, v, self.t[self.i - 5:self.i + 5])

    def type_path(self):
        # consumes things like []wildcard.Wildcard, map[string]string, *unstructured.Unstructured, metav1.X
        s = ""
        while True:
            k, v = self.peek()
            if v == "[":
                self.next()
                if self.accept("]"):
                    s += "[]"
                else:
                    inner = self.type_path()
                    self.expect("]")
                    s += "[" + inner + "]"
            elif v == "*":
                self.next()
                s += "*"
            elif k == "id":
                self.next()
                s += v
                while self.peek()[1] == "." and self.peek(1)[0] == "id":
                    self.next()
                    s += "." + self.next()[1]
                if s.endswith("map") or s == "map":
                    continue
                return s
            else:
                return s

    def elems(self):
        items = []
        self.expect("{")
        while not self.accept("}"):
            v = self.expr()
            if self.accept(":"):
                items.append((v, self.expr()))
            else:
                items.append((None, v))
            self.accept(",")
        return items

    def expr(self):
        k, v = self.peek()
        if v == "&":
            self.next()
            return self.expr()
        if v == "{":
            return ("lit", "", self.elems())
        if k == "str":
            self.next()
            return ("str", json.loads(v))
        if k == "raw":
            self.next()
            return ("str", v[1:-1])
        if k == "num":
            self.next()
            return ("num", int(v))
        if v == "func":
            self.next()
            depth = 0
            # skip signature
            while self.peek()[1] != "{":
                self.next()
            start = self.i
            while True:
                _, x = self.next()
                if x == "{":
                    depth += 1
                elif x == "}":
                    depth -= 1
                    if depth == 0:
                        break
            body = self.t[start:self.i]
            labels = None
            for j, (_, x) in enumerate(body):
                if x == "SetLabels":
                    sub = P(body[j + 2:])
                    labels = sub.expr()
            return ("func", labels)
        if v == "[" or v == "map" or v == "*":
            tp = self.type_path()
            return ("lit", tp, self.elems())
        if k == "id":
            tp = self.type_path()
            if self.peek()[1] == "(":
                self.next()
                args = []
                while not self.accept(")"):
                    args.append(self.expr())
                    self.accept(",")
                return ("call", tp, args)
            if self.peek()[1] == "{":
                return ("lit", tp, self.elems())
            return ("ident", tp)
        raise SystemExit(f"unexpected token {k} {v!r} near {self.t[self.i-5:self.i+5]}")


CONST = {
    "apiextensionsv1.ClusterScoped": "Cluster", "apiextensionsv1.NamespaceScoped": "Namespaced",
    "types.SourceTypeOriginal": "Original", "types.SourceTypeGenerated": "Generated", "types.SourceTypeAll": "All",
    "types.SourceTypeDefault": "All",
    "metav1.LabelSelectorOpIn": "In", "metav1.LabelSelectorOpNotIn": "NotIn",
    "metav1.LabelSelectorOpExists": "Exists", "metav1.LabelSelectorOpDoesNotExist": "DoesNotExist",
    "Wildcard": "*", "nil": None, "true": True, "false": False,
}


def val(n):
    k = n[0]
    if k == "str" or k == "num":
        return n[1]
    if k == "ident":
        if n[1] in CONST:
            return CONST[n[1]]
        raise SystemExit(f"unknown constant {n[1]}")
    if k == "call":
        if n[1] in ("string", "types.SourceType", "wildcard.Wildcard", "Wildcard"):
            return val(n[2][0])
        raise SystemExit(f"unknown call {n[1]}")
    if k == "lit":
        tp = n[1]
        if tp.startswith("map"):
            return {val(a): val(b) for a, b in n[2]}
        if tp.startswith("[]"):
            return [val(b) if b[0] != "lit" or b[1] else struct(b, tp[2:]) for _, b in n[2]]
        return struct(n, tp)
    raise SystemExit(f"cannot evaluate {n}")


FIELD = {"Kinds": "kinds", "APIGroups": "apiGroups", "Scope": "scope", "Namespaces": "namespaces",
         "ExcludedNamespaces": "excludedNamespaces", "LabelSelector": "labelSelector",
         "NamespaceSelector": "namespaceSelector", "Name": "name", "Source": "source",
         "MatchLabels": "matchLabels", "MatchExpressions": "matchExpressions", "Key": "key",
         "Operator": "operator", "Values": "values", "Labels": "labels", "ObjectMeta": "metadata"}


def struct(n, tp):
    out = {}
    for a, b in n[2]:
        assert a is not None and a[0] == "ident", n
        out[FIELD[a[1]]] = val(b)
    return out


def make_object(n):
    if n[0] == "ident" and n[1] == "nil":
        return None
    assert n[0] == "call", n
    fn, args = n[1], n[2]
    labels = None
    if fn in ("makeObject", "makeObjectWithGenerateName"):
        gvk = {a[1]: val(b) for a, b in args[0][2]}
        group, version, kind = gvk.get("Group", ""), gvk.get("Version", ""), gvk.get("Kind", "")
        api = f"{group}/{version}" if group else version
        ns, name = val(args[1]), val(args[2])
        md = {}
        if ns:
            md["namespace"] = ns
        if name:
            md["generateName" if fn == "makeObjectWithGenerateName" else "name"] = name
        obj = {"apiVersion": api, "kind": kind, "metadata": md}
        rest = args[3:]
    elif fn == "makeNamespace":
        obj = {"apiVersion": "v1", "kind": "Namespace",
               "metadata": {"name": val(args[0]), "creationTimestamp": None}, "spec": {}, "status": {}}
        rest = args[1:]
    else:
        raise SystemExit(f"unknown object helper {fn}")
    for f in rest:
        assert f[0] == "func"
        if f[1] is not None:
            labels = val(f[1])
    if labels is not None:
        obj["metadata"]["labels"] = labels
    return obj


def make_ns(n):
    if n is None or (n[0] == "ident" and n[1] == "nil"):
        return None
    d = val(n)
    ns = {"apiVersion": "v1", "kind": "Namespace", "metadata": d.get("metadata", {})}
    return ns


def table(src, func_name, var):
    i = src.index("func " + func_name + "(")
    j = src.index(var, i)
    j = src.index("{", src.index("}", j))  # after the anonymous struct type
    toks = lex(src[j:])
    p = P(toks)
    return p.elems()


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    src = open(f"{REF}/pkg/mutation/match/match_test.go").read()
    out = []
    for _, row in table(src, "TestMatch", "table := []struct"):
        f = {a[1]: b for a, b in row[2]}
        out.append({
            "source_test": "pkg/mutation/match/match_test.go:TestMatch",
            "name": val(f["name"]),
            "object": make_object(f["object"]),
            "match": val(f["matcher"]),
            "namespace": make_ns(f.get("namespace")),
            "source": val(f["source"]) if "source" in f else "",
            "wantMatch": val(f["wantMatch"]) if "wantMatch" in f else False,
            "wantErr": "wantErr" in f and val(f["wantErr"]) is not None if f.get("wantErr", ("ident", "nil"))[1] in CONST else True,
        })
    for _, row in table(src, "Test_namesMatch", "tests := []struct"):
        f = {a[1]: b for a, b in row[2]}
        args = {a[1]: b for a, b in f["args"][2]}
        tgt = {a[1]: b for a, b in args["target"][2]}
        out.append({
            "source_test": "pkg/mutation/match/match_test.go:Test_namesMatch",
            "name": val(f["name"]),
            "object": make_object(tgt["Object"]),
            "match": val(args["match"]),
            "namespace": make_ns(tgt.get("Namespace")),
            "source": val(tgt["Source"]) if "Source" in tgt else "",
            "wantMatch": val(f["want"]),
            "wantErr": val(f["wantErr"]) if "wantErr" in f else False,
            "only": "name",
        })
    json.dump(out, open(f"{HERE}/match_vectors.json", "w"), indent=1)
    print(f"match_vectors.json: {len(out)} vectors")

    wsrc = open(f"{REF}/pkg/wildcard/wildcard_test.go").read()
    wout = []
    for fn, var, kind in (("TestMatches", "tcs := []struct", "matches"),
                          ("TestWildcard_MatchesGenerateName", "tcs := []struct", "generateName")):
        try:
            rows = table(wsrc, fn, var)
        except ValueError:
            continue
        for _, row in rows:
            f = {a[1]: val(b) for a, b in row[2]}
            wout.append({"fn": kind, **f})
    json.dump(wout, open(f"{HERE}/wildcard_vectors.json", "w"), indent=1)
    print(f"wildcard_vectors.json: {len(wout)} vectors")


if __name__ == "__main__":
    main()
