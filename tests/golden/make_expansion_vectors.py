#!/usr/bin/env python
"""Extracts the reference's expansion vectors that involve no mutators (the mutation system is outside this engine's scope):
pkg/expansion/system_test.go TestExpand table entries + the YAML fixtures they name (pkg/expansion/fixtures/fixtures.go), and the
gator expansion manifests with the messages test/gator/test/test.bats asserts.  Authoring container only; data, no code.

    python tests/golden/make_expansion_vectors.py   ->  tests/golden/expansion_vectors.json
"""
import json
import os
import re

import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

fx_src = open(os.path.join(REF, "pkg/expansion/fixtures/fixtures.go")).read()
FIX = {m.group(1): yaml.safe_load(m.group(2)) for m in re.finditer(r"\n\t(\w+) = `\n(.*?)`", fx_src, re.S)}

test_src = open(os.path.join(REF, "pkg/expansion/system_test.go")).read()
start = test_src.index("func TestExpand(t *testing.T)")
end = test_src.index("\n\tfor _, tc := range tests", start)
table = test_src[start:end]
line0 = test_src[:start].count("\n") + 1
cases = []
# entries start with "\n\t\t{\n\t\t\tname:" and end with "\n\t\t},"
for m in re.finditer(r"\n\t\t\{\n(\t\t\tname:.*?)\n\t\t\},", table, re.S):
    body = m.group(1)
    name = re.search(r'name:\s+"([^"]*)"', body).group(1)
    if re.search(r"mutators:\s*\[\]types\.Mutator\{\s*\n", body):
        continue                                            # a non-empty mutator list: needs the mutation system
    gen = re.search(r"generator:\s+fixtures\.LoadFixture\(fixtures\.(\w+), t\)", body).group(1)
    ns = re.search(r'ns:\s+&corev1\.Namespace\{ObjectMeta: metav1\.ObjectMeta\{Name: "([^"]*)"\}\}', body)
    tmpls = re.findall(r"fixtures\.LoadTemplate\(fixtures\.(\w+), t\)", body)
    want = [{"obj": FIX[o], "enforcementAction": a, "templateName": t} for o, a, t in
            re.findall(r'\{Obj: fixtures\.LoadFixture\(fixtures\.(\w+), t\), EnforcementAction: "([^"]*)", TemplateName: "([^"]*)"\}', body)]
    cases.append({"name": name, "source": "pkg/expansion/system_test.go:%d" % (line0 + table[:m.start()].count("\n") + 1),
                  "generator": FIX[gen], "ns": ns.group(1) if ns else None, "templates": [FIX[t] for t in tmpls], "want": want,
                  "expectErr": bool(re.search(r"expectErr:\s+true", body))})

# gator: expansion with a namespace selector (test/gator/test/test.bats:268-289)
gdir = "test/gator/test/fixtures/manifests/expansion"
docs = [d for d in yaml.safe_load_all(open(os.path.join(REF, gdir, "expansion-w-ns-selector.yaml"))) if d]
nsdocs = [d for d in yaml.safe_load_all(open(os.path.join(REF, gdir, "ns.yaml"))) if d]
bats = open(os.path.join(REF, "test/gator/test/test.bats")).read()
m1 = re.search(r'want_msg="(Implied by expand-deployments\] unable to match constraints[^"]*)"', bats)
m2 = re.search(r'want_msg="(\[Implied by expand-deployments\] All pods must have[^"]*)"', bats)
gator = {"source": "test/gator/test/test.bats:268-289", "docs": docs, "ns_docs": nsdocs,
         "without_ns_substring": m1.group(1), "with_ns_substring": m2.group(1).replace("\\`", "`")}



# ---- Go composite literals of map[string]interface{} / []interface{} / strings / ints -> Python values
def go_literal(text):
    pos = 0
    text = re.sub(r"//[^\n]*", "", text)

    def ws():
        nonlocal pos
        while pos < len(text) and text[pos] in " \t\n,":
            pos += 1

    def value():
        nonlocal pos
        ws()
        if text.startswith("map[string]interface{}{", pos):
            pos += len("map[string]interface{}{")
            out = {}
            while True:
                ws()
                if text[pos] == "}":
                    pos += 1
                    return out
                k = value()
                ws()
                assert text[pos] == ":", text[pos:pos + 30]
                pos += 1
                out[k] = value()
        if text.startswith("[]interface{}{", pos):
            pos += len("[]interface{}{")
            out = []
            while True:
                ws()
                if text[pos] == "}":
                    pos += 1
                    return out
                out.append(value())
        if text[pos] == '"':
            m = re.compile(r'"((?:[^"\\]|\\.)*)"').match(text, pos)
            pos = m.end()
            return json.loads('"' + m.group(1) + '"')
        m = re.compile(r"-?\d+").match(text, pos)
        assert m, text[pos:pos + 40]
        pos = m.end()
        return int(m.group(0))
    return value()


def table_entries(func, until="\n\tfor _, tc := range tests"):
    a = test_src.index("func %s(t *testing.T)" % func)
    b = test_src.index(until, a)
    tab = test_src[a:b]
    l0 = test_src[:a].count("\n") + 1
    for m in re.finditer(r"\n\t\t\{\n(\t\t\tname:.*?)\n\t\t\},(?=\n\t\t\{|\n\t\})", tab, re.S):
        yield m.group(1), "pkg/expansion/system_test.go:%d" % (l0 + tab[:m.start()].count("\n") + 1)


def strs(body, field):
    m = re.search(field + r":\s+\[\]string\{([^}]*)\}", body)
    return re.findall(r'"([^"]*)"', m.group(1)) if m else None


def gen_gvk(body, field):
    m = re.search(field + r":\s+expansionunversioned\.GeneratedGVK\{(.*?)\}", body, re.S)
    if not m:
        return None
    return {k.lower(): v for k, v in re.findall(r'(Group|Version|Kind):\s+"([^"]*)"', m.group(1))}


def template_doc(name, apply, source, gvk):
    spec = {}
    if apply:
        spec["applyTo"] = apply
    if source is not None:
        spec["templateSource"] = source
    if gvk is not None:
        spec["generatedGVK"] = gvk
    return {"apiVersion": "expansion.gatekeeper.sh/v1beta1", "kind": "ExpansionTemplate", "metadata": {"name": name}, "spec": spec}


# TestValidateTemplate (system_test.go:311-424): template -> the substring its error must contain (None: valid)
validate = []
for body, src in table_entries("TestValidateTemplate"):
    name = re.search(r'name:\s+"([^"]*)"', body).group(1)
    want = re.search(r'errFn:\s+matchErr\("([^"]*)"\)', body)
    if "fixtures.TestTemplate(" in body:        # TestTemplate("foo", 1, 2) -- fixtures/load.go:127-142
        nm, a, g = re.search(r'TestTemplate\("([^"]*)", (\d+), (\d+)\)', body).groups()
        doc = template_doc(nm, [{"groups": ["group" + a], "versions": ["v" + a], "kinds": ["kind" + a]}], "spec.template",
                           {"group": "group" + g, "version": "v" + g, "kind": "kind" + g})
    else:
        tn = re.search(r'\n\t\t\t\tName:\s+"([^"]*)"', body)
        apply = None
        if "Apply:" in body:
            apply = [{"groups": strs(body, "Groups"), "kinds": strs(body, "Kinds"), "versions": strs(body, "Versions")}]
        srcf = re.search(r'\n\t\t\t\tSource:\s+"([^"]*)"', body)
        doc = template_doc(tn.group(1) if tn else "", apply, srcf.group(1) if srcf else None, gen_gvk(body, "GenGVK"))
    validate.append({"name": name, "source": src, "template": doc, "errSubstr": want.group(1) if want else None})

# TestExpandResource (system_test.go:426-660): (object, namespace, template) -> resultant | error substring
expand_resource = []
for body, src in table_entries("TestExpandResource", until="\n\tfor _, tc := range tests"):
    name = re.search(r'name:\s+"([^"]*)"', body).group(1)
    m = re.search(r"\n\t\t\tobj:\s+fixtures\.LoadFixture\(fixtures\.(\w+), t\)", body)
    if m:
        obj = FIX[m.group(1)]
    else:
        a = body.index("Object: ", body.index("\n\t\t\tobj:")) + len("Object: ")
        obj = go_literal(body[a:])
    ns = re.search(r'\n\t\t\tns:\s+&corev1\.Namespace\{ObjectMeta: metav1\.ObjectMeta\{Name: "([^"]*)"\}\}', body)
    tsrc = re.search(r'TemplateSource:\s+"([^"]*)"', body).group(1)
    tname = re.search(r'ObjectMeta: metav1\.ObjectMeta\{Name: "([^"]*)"\},\n\t\t\t\tSpec', body).group(1)
    tdoc = template_doc(tname, None, tsrc, gen_gvk(body, "GeneratedGVK") or {})
    want = None
    m = re.search(r"\n\t\t\twant:\s+fixtures\.LoadFixture\(fixtures\.(\w+), t\)", body)
    if m:
        want = FIX[m.group(1)]
    elif "\n\t\t\twant: &unstructured.Unstructured{" in body:
        a = body.index("Object: ", body.index("\n\t\t\twant:")) + len("Object: ")
        want = go_literal(body[a:])
    es = re.search(r'errSubstr:\s+"([^"]*)"', body)
    expand_resource.append({"name": name, "source": src, "obj": obj, "ns": ns.group(1) if ns else None, "template": tdoc, "want": want,
                            "errSubstr": es.group(1) if es else None})

# TestDB (db_test.go:27-647): sequences of template upserts / removals -> which stored templates are set aside as part of an expansion
# cycle (hasConflicts), and which upsert must report "template forms expansion cycle"
def test_template(name, a, g):     # fixtures.TestTemplate -- fixtures/load.go:127-142
    return template_doc(name, [{"groups": ["group%s" % a], "versions": ["v%s" % a], "kinds": ["kind%s" % a]}], "spec.template",
                        {"group": "group%s" % g, "version": "v%s" % g, "kind": "kind%s" % g})


TEMP_MULT_APPLY = template_doc("t2", [{"groups": ["group1"], "versions": ["v1"], "kinds": ["kind1"]},       # fixtures.TempMultApply -- load.go:144-166
                                      {"groups": ["group11"], "versions": ["v11", "v22"], "kinds": ["kind11"]}], "spec.template",
                               {"group": "group2", "version": "v2", "kind": "kind2"})
db_src = open(os.path.join(REF, "pkg/expansion/db_test.go")).read()
a0 = db_src.index("func TestDB(t *testing.T)")
b0 = db_src.index("\n\tfor _, tc := range tests", a0)
db_tab = db_src[a0:b0]
db_line0 = db_src[:a0].count("\n") + 1
TREF = r'\*?fixtures\.(?:TestTemplate\("([^"]*)", (\d+), (\d+)\)|(TempMultApply)\(\))'
db_cases = []
for m in re.finditer(r"\n\t\t\{\n\t\t\tname: \"([^\"]*)\",\n(.*?)(?=\n\t\t\{\n\t\t\tname: |\Z)", db_tab, re.S):
    body = m.group(2)
    ops_txt = body[body.index("ops: []templateOperation{"):body.index("wantStore:") if "wantStore:" in body else len(body)]
    ops = []
    for o in re.finditer(r"op:\s+(addOp|rmOp),\s*\n\s*template:\s+" + TREF + r",(?:\s*\n\s*wantErr:\s+(true|false),)?", ops_txt):
        doc = TEMP_MULT_APPLY if o.group(5) else test_template(o.group(2), o.group(3), o.group(4))
        ops.append({"op": "add" if o.group(1) == "addOp" else "remove", "template": doc, "wantErr": o.group(6) == "true"})
    assert len(ops) == ops_txt.count("op: "), (m.group(1), len(ops), ops_txt.count("op: "))
    want = {}
    if "wantStore:" in body:
        st = body[body.index("wantStore:"):body.index("wantMatchers:") if "wantMatchers:" in body else len(body)]
        for w in re.finditer(r"keyForTemplate\(" + TREF + r"\): \{\s*\n\s*template:[^\n]*\n\s*hasConflicts:\s+(true|false),", st):
            want["t2" if w.group(4) else w.group(1)] = w.group(5) == "true"
        assert len(want) == st.count("hasConflicts:"), (m.group(1), want)
    db_cases.append({"name": m.group(1), "source": "pkg/expansion/db_test.go:%d" % (db_line0 + db_tab[:m.start()].count("\n") + 1), "ops": ops, "want": want})

# TestApplyTo (pkg/mutation/match/match_test.go:717-845): ApplyTo.Matches, the relation that selects an expansion template for a GVK
mt_src = open(os.path.join(REF, "pkg/mutation/match/match_test.go")).read()
a1 = mt_src.index("func TestApplyTo(t *testing.T)")
b1 = mt_src.index("\n\tfor _, tc := range table", a1)
at_tab = mt_src[a1:b1]
at_line0 = mt_src[:a1].count("\n") + 1
apply_to = []
for m in re.finditer(r"\n\t\t\{\n\t\t\tname: \"([^\"]*)\",\n(.*?)\n\t\t\},(?=\n\t\t\{|\n\t\})", at_tab, re.S):
    body = m.group(2)
    g = re.search(r'gvk:\s+schema\.GroupVersionKind\{Group: "([^"]*)", Version: "([^"]*)", Kind: "([^"]*)"\}', body)
    entries = [{"groups": re.findall(r'"([^"]*)"', e.group(1)), "versions": re.findall(r'"([^"]*)"', e.group(2)), "kinds": re.findall(r'"([^"]*)"', e.group(3))}
               for e in re.finditer(r"Groups:\s+\[\]string\{([^}]*)\},\s*Versions:\s+\[\]string\{([^}]*)\},\s*Kinds:\s+\[\]string\{([^}]*)\}", body)]
    assert entries and g, m.group(1)
    apply_to.append({"name": m.group(1), "source": "pkg/mutation/match/match_test.go:%d" % (at_line0 + at_tab[:m.start()].count("\n") + 1),
                     "gvk": list(g.groups()), "applyTo": entries, "wantApply": re.search(r"wantApply:\s+(true|false)", body).group(1) == "true"})

with open(os.path.join(HERE, "expansion_vectors.json"), "w") as f:
    json.dump({"expand": cases, "gator": gator, "validate": validate, "expand_resource": expand_resource, "db": db_cases, "apply_to": apply_to}, f, indent=1, sort_keys=True)
print("TestApplyTo:", [(c["name"], c["wantApply"], len(c["applyTo"])) for c in apply_to])
print("TestDB:", [(c["name"], len(c["ops"]), sum(o["wantErr"] for o in c["ops"]), c["want"]) for c in db_cases])
print("TestValidateTemplate:", [(v["name"], v["errSubstr"]) for v in validate])
print("TestExpandResource:", [(v["name"], v["errSubstr"], v["want"] is not None) for v in expand_resource])
print("TestExpand cases without mutators:", len(cases))
for c in cases:
    print("  ", c["source"], c["name"], "->", len(c["want"]), "err" if c["expectErr"] else "")
print("gator docs:", [d.get("kind") for d in docs], [d.get("kind") for d in nsdocs])
