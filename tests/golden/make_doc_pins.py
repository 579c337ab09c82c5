#!/usr/bin/env python
"""Message pins the reference holds outside its test tables: the documentation quotes exact violation messages (arrays, objects
and strings through OPA's sprintf("%v")), and test/gator/test/fixtures/manifests/with-policies/*-2.yaml are the AllowedRepos
manifests of the gator suite.  This script reads them from /root/reference (authoring container only) and writes
tests/golden/doc_pins.json; the quoted messages are located in the docs by regular expression so that a change there breaks the
script, not silently the pin.  Data only -- no reference code.

    python tests/golden/make_doc_pins.py
"""
import json
import os
import re
import sys

import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_all(path):
    with open(os.path.join(REF, path)) as f:
        return [d for d in yaml.safe_load_all(f) if d]


def find(path, pattern):
    with open(os.path.join(REF, path)) as f:
        for no, line in enumerate(f, 1):
            m = re.search(pattern, line)
            if m:
                return m.group(1), "%s:%d" % (path, no)
    raise SystemExit("pattern %r not found in %s" % (pattern, path))


def ct(path):
    return [d for d in load_all(path) if d.get("kind") == "ConstraintTemplate"][0]


def constraint(kind, name, params=None, action=None, match=None):
    c = {"apiVersion": "constraints.gatekeeper.sh/v1beta1", "kind": kind, "metadata": {"name": name}, "spec": {}}
    if params is not None:
        c["spec"]["parameters"] = params
    if action:
        c["spec"]["enforcementAction"] = action
    c["spec"]["match"] = match or {"kinds": [{"apiGroups": [""], "kinds": ["Pod"]}]}
    return c


def pod(name, containers):
    return {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": name, "namespace": "default"}, "spec": {"containers": containers}}


cases = []
# 1. website/docs/violations.md: %v of an ARRAY parameter
msg, src = find("website/docs/violations.md", r"Warning: \[[a-z-]+\] (container <nginx> has an invalid image repo <nginx>, allowed repos are \[\"openpolicyagent\"\])")
cases.append({"name": "allowedrepos-doc-warn", "source": src, "ep": "validation.gatekeeper.sh",
              "docs": [ct("demo/agilebank/templates/k8sallowedrepos_template.yaml"),
                       constraint("K8sAllowedRepos", "repo-is-openpolicyagent", {"repos": ["openpolicyagent"]}, action="warn"),
                       pod("pause", [{"name": "nginx", "image": "nginx"}])],
              "expect": [{"constraint": "K8sAllowedRepos/repo-is-openpolicyagent", "msg": msg, "action": "warn"}]})
# 2. website/docs/workload-resources.md: %v of an OBJECT
msg, src = find("website/docs/workload-resources.md", r"denied the request: \[psp-privileged-container\] (Privileged container is not allowed: nginx, securityContext: \{\"privileged\": true\})$")
cases.append({"name": "psp-privileged-doc", "source": src, "ep": "validation.gatekeeper.sh",
              "docs": [ct("pkg/webhook/testdata/psp-all-violations/psp-templates/privileged-containers-template.yaml"),
                       constraint("K8sPSPPrivilegedContainer", "psp-privileged-container"),
                       pod("i-wont-be-blocked-755547df65-x", [{"name": "nginx", "image": "nginx", "securityContext": {"privileged": True}}])],
              "expect": [{"constraint": "K8sPSPPrivilegedContainer/psp-privileged-container", "msg": msg, "action": "deny"}]})
# 3. website/docs/audit.md: the audit log line of K8sContainerLimits
msg, src = find("website/docs/audit.md", r"\"msg\": \"(container <kube-scheduler> has no resource limits)\"")
cases.append({"name": "containerlimits-doc-audit", "source": src, "ep": "audit.gatekeeper.sh",
              "docs": [ct("demo/agilebank/templates/k8scontainterlimits_template.yaml"),
                       constraint("K8sContainerLimits", "container-must-have-limits", {"cpu": "200m", "memory": "1Gi"}),
                       pod("kube-scheduler-kind-control-plane", [{"name": "kube-scheduler", "image": "k8s.gcr.io/kube-scheduler:v1.21.1"}])],
              "expect": [{"constraint": "K8sContainerLimits/container-must-have-limits", "msg": msg, "action": "deny"}]})
# 4. the AllowedRepos manifests of the gator suite (test/gator/test/test.bats runs them: violations <=> non-zero exit)
for fn, violates in (("with-violations-2.yaml", True), ("no-violations-2.yaml", False)):
    path = "test/gator/test/fixtures/manifests/with-policies/" + fn
    docs = load_all(path)
    exp = []
    if violates:
        repos = [d for d in docs if d.get("kind") == "K8sAllowedRepos"][0]["spec"]["parameters"]["repos"]
        for d in docs:
            if d.get("kind") == "Pod":
                for c in d["spec"]["containers"]:
                    if not any(c["image"].startswith(r) for r in repos):
                        # the message format is the one violations.md quotes (an array prints as ["a", "b"])
                        exp.append({"constraint": "K8sAllowedRepos/" + [x for x in docs if x.get("kind") == "K8sAllowedRepos"][0]["metadata"]["name"],
                                    "msg": "container <%s> has an invalid image repo <%s>, allowed repos are [%s]" % (
                                        c["name"], c["image"], ", ".join('"%s"' % r for r in repos)), "action": "deny"})
    cases.append({"name": "gator-" + fn.replace(".yaml", ""), "source": path, "ep": "gator.gatekeeper.sh", "docs": docs, "expect": exp,
                  "expect_any": violates})

with open(os.path.join(HERE, "doc_pins.json"), "w") as f:
    json.dump(cases, f, indent=1, sort_keys=True)
print("wrote %d cases" % len(cases))
for c in cases:
    print(" ", c["name"], c["source"], len(c["expect"]))
