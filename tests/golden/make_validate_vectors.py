#!/usr/bin/env python
"""tests/golden/validate_constraint_vectors.json: the 11 cases of TestValidateConstraint (pkg/target/target_test.go:42-399) --
a constraint document and whether K8sValidationTarget.ValidateConstraint (pkg/target/target.go:178-214) must reject it
(selector of the wrong type, matchLabels of the wrong type, unknown matchExpressions operator; labelSelector and
namespaceSelector).  Data only.  Run in the authoring container: python tests/golden/make_validate_vectors.py"""
import json
import os
import re
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    src = open(f"{REF}/pkg/target/target_test.go").read()
    body = src[src.index("func TestValidateConstraint"):src.index("func TestProcessData")]
    cases = re.findall(r'Name:\s*"([^"]+)",\s*Constraint:\s*`(.*?)`,\s*ErrorExpected:\s*(true|false)', body, re.S)
    out = {"source": "pkg/target/target_test.go:42-399 (TestValidateConstraint)",
           "cases": [{"name": n, "constraint": json.loads(c), "error_expected": e == "true"} for n, c, e in cases]}
    assert len(out["cases"]) == 11
    with open(os.path.join(HERE, "validate_constraint_vectors.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote validate_constraint_vectors.json:", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
