#!/usr/bin/env python
"""Collects the reference's own policy fixtures into tests/golden/*.json so that parity tests and the
benchmark can run on the GPU box, where /root/reference does not exist.  Data only (ConstraintTemplates,
Constraints, example objects, golden messages asserted by the reference's tests) -- no reference code.

Run in the authoring container:  python tests/golden/make_fixtures.py

Outputs
  templates.json     every in-tree ConstraintTemplate used by configs 1-5, with source path
  psp_suite.json     pkg/webhook/testdata/psp-all-violations (5 templates, 5 constraints, 5 pods) -- the
                     BenchmarkValidationHandler fixture set (pkg/webhook/policy_benchmark_test.go:264-271)
  gator_cases.json   gator test manifests + the exact messages the reference's bats tests assert
                     (test/gator/test/test.bats:73,95,114,165-202,233,249), verify suite (test/gator/verify/suite.yaml)
  oracle_outputs.json violation sets produced by the oracle for the above (regression pin for the oracle itself)
"""
import glob
import json
import os
import sys

import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def load_all(path):
    with open(path) as f:
        return [d for d in yaml.safe_load_all(f) if d]


def rego_of(ct):
    tgt = ct["spec"]["targets"][0]
    src = tgt.get("rego")
    if not src:
        for c in tgt.get("code") or []:
            if c.get("engine") == "Rego":
                src = c["source"]["rego"]
    return src


TEMPLATE_FILES = {
    "requiredlabels_basic": "test/gator/bench/basic/template.yaml",
    "allowedrepos_prefixmatch": "test/gator/bench/both/template.yaml",
    "allowedrepos": "demo/agilebank/templates/k8sallowedrepos_template.yaml",
    "containerlimits": "demo/agilebank/templates/k8scontainterlimits_template.yaml",
    "requiredprobes": "demo/agilebank/templates/k8srequiredprobes_template.yaml",
    "bannedimagetags": "demo/agilebank/remediation/k8sbannedimagetags_template.yaml",
    "requiredlabels_agilebank": "demo/agilebank/templates/k8srequiredlabels_template.yaml",
    "requiredlabels_regov1": "test/bats/tests/templates/k8srequiredlabels_template_regov1.yaml",
    "requiredlabels_example": "example/templates/k8srequiredlabels_template.yaml",
    "fooischeck": "test/gator/verify/template.yaml",
    "namespacelabelcheck": "test/bats/tests/templates/k8snamespacelabelcheck_template_rego.yaml",
    "psp_privileged": "pkg/webhook/testdata/psp-all-violations/psp-templates/privileged-containers-template.yaml",
    "psp_hostnamespace": "pkg/webhook/testdata/psp-all-violations/psp-templates/host-namespace-template.yaml",
    "psp_hostnetworkports": "pkg/webhook/testdata/psp-all-violations/psp-templates/host-network-ports-template.yaml",
    "psp_volumetypes": "pkg/webhook/testdata/psp-all-violations/psp-templates/volume-template.yaml",
    "psp_hostfilesystem": "pkg/webhook/testdata/psp-all-violations/psp-templates/host-filesystem-template.yaml",
}


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    templates = {}
    for name, rel in TEMPLATE_FILES.items():
        ct = load_all(f"{REF}/{rel}")[0]
        templates[name] = {"kind": ct["spec"]["crd"]["spec"]["names"]["kind"], "rego": rego_of(ct), "source": rel}
    # Go-string fixtures (pkg/gator/fixtures/fixtures.go): templates embedded as YAML in Go raw strings
    import re
    src = open(f"{REF}/pkg/gator/fixtures/fixtures.go").read()
    for m in re.finditer(r"\t(\w+)\s*=\s*`\n(.*?)`", src, re.S):
        try:
            doc = yaml.safe_load(m.group(2))
        except yaml.YAMLError:
            continue
        if isinstance(doc, dict) and doc.get("kind") == "ConstraintTemplate" and isinstance(doc.get("spec"), dict):
            try:
                r = rego_of(doc)
            except (KeyError, TypeError):
                continue
            if r:
                templates["fixtures_" + m.group(1)] = {"kind": doc["spec"]["crd"]["spec"]["names"]["kind"], "rego": r,
                                                       "source": "pkg/gator/fixtures/fixtures.go:" + m.group(1)}
    json.dump(templates, open(f"{HERE}/templates.json", "w"), indent=1)
    print("templates.json:", len(templates))

    base = f"{REF}/pkg/webhook/testdata/psp-all-violations"
    psp = {"templates": [], "constraints": [], "pods": []}
    for p in sorted(glob.glob(base + "/psp-templates/*.yaml")):
        ct = load_all(p)[0]
        psp["templates"].append({"kind": ct["spec"]["crd"]["spec"]["names"]["kind"], "rego": rego_of(ct)})
    for p in sorted(glob.glob(base + "/psp-constraints/*.yaml")):
        psp["constraints"] += load_all(p)
    for p in sorted(glob.glob(base + "/psp-pods/*.yaml")):
        psp["pods"] += load_all(p)
    json.dump(psp, open(f"{HERE}/psp_suite.json", "w"), indent=1)
    print("psp_suite.json:", len(psp["pods"]), "pods")

    g = f"{REF}/test/gator/test/fixtures"
    cases = []

    def case(name, files, must_contain, source):
        docs = []
        for f in files:
            paths = sorted(glob.glob(f"{g}/{f}/*.yaml")) if os.path.isdir(f"{g}/{f}") else [f"{g}/{f}"]
            for p in paths:
                docs += load_all(p)
        cases.append({"name": name, "docs": docs, "must_contain": must_contain, "source": source})

    probe_msg = "Container <tomcat> in your <Pod> <test-pod1> has no <readinessProbe>"
    case("with-policies/with-violations", ["manifests/with-policies/with-violations.yaml"], [probe_msg], "test/gator/test/test.bats:73")
    case("with-policies/no-violations", ["manifests/with-policies/no-violations.yaml"], [], "test/gator/test/test.bats:76-82")
    case("with-policies/rego-v1", ["manifests/with-policies/with-violations-rego-v1.yaml"], None, "test/gator/test/test.bats:88-90")
    case("default policies + with-violations", ["policies/default", "manifests/no-policies/with-violations.yaml"], [probe_msg],
         "test/gator/test/test.bats:92-96")
    case("default policies + no-violations", ["policies/default", "manifests/no-policies/no-violations.yaml"], [],
         "test/gator/test/test.bats:117-123")
    json.dump(cases, open(f"{HERE}/gator_cases.json", "w"), indent=1)
    print("gator_cases.json:", len(cases))

    # verify suite (K8sFooIs): allow/deny pinned by test/gator/verify/suite.yaml
    v = f"{REF}/test/gator/verify"
    verify = {"template": load_all(f"{v}/template.yaml")[0], "files": {}}
    for p in sorted(glob.glob(f"{v}/*.yaml")):
        verify["files"][os.path.basename(p)] = load_all(p)
    json.dump(verify, open(f"{HERE}/verify_suite.json", "w"), indent=1)
    print("verify_suite.json:", len(verify["files"]), "files")


if __name__ == "__main__":
    main()
