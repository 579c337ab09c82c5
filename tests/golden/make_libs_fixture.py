#!/usr/bin/env python
"""tests/golden/libs_template.json: the reference's one in-tree ConstraintTemplate that uses `spec.targets[].libs`
(test/bats/tests/templates/k8scontainterlimits_template.yaml: `package lib.helpers`, imported as data.lib.helpers), its
constraint and the two Pods the bats test applies (test/bats/test.bats:268-279: bad/opa_no_limits.yaml is denied,
good/opa.yaml is admitted).  Data only.  Run in the authoring container: python tests/golden/make_libs_fixture.py"""
import json
import os
import sys

import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(rel):
    with open(f"{REF}/{rel}") as f:
        return [d for d in yaml.safe_load_all(f) if d]


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    out = {
        "source": "test/bats/tests/templates/k8scontainterlimits_template.yaml, test/bats/test.bats:268-279",
        "template": load("test/bats/tests/templates/k8scontainterlimits_template.yaml")[0],
        "constraint": load("test/bats/tests/constraints/containers_must_be_limited.yaml")[0],
        "denied": load("test/bats/tests/bad/opa_no_limits.yaml")[0],
        "admitted": load("test/bats/tests/good/opa.yaml")[0],
    }
    with open(os.path.join(HERE, "libs_template.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote libs_template.json")


if __name__ == "__main__":
    main()
