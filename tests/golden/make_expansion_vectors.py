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

with open(os.path.join(HERE, "expansion_vectors.json"), "w") as f:
    json.dump({"expand": cases, "gator": gator}, f, indent=1, sort_keys=True)
print("TestExpand cases without mutators:", len(cases))
for c in cases:
    print("  ", c["source"], c["name"], "->", len(c["want"]), "err" if c["expectErr"] else "")
print("gator docs:", [d.get("kind") for d in docs], [d.get("kind") for d in nsdocs])
