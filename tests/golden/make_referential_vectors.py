#!/usr/bin/env python
"""Referential-constraint fixtures taken from the reference tree (run where /root/reference exists):

  * test/bats/tests/templates/k8suniquelabel_template.yaml + constraints/all_cm_gatekeeper_label_unique.yaml with the
    objects of the "unique labels test" (test/bats/test.bats:295-303): good/no_dupe_cm.yaml is synced, then
    bad/no_dupe_cm_2.yaml must be denied.
  * the K8sUniqueServiceSelector rows of pkg/gator/test/test_test.go live in gator_test_table.json already.

Writes tests/golden/referential_vectors.json."""
import json
import os
import yaml

REF = "/root/reference/test/bats/tests"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(rel):
    with open(os.path.join(REF, rel)) as f:
        return [d for d in yaml.safe_load_all(f) if d]


out = {
    "uniquelabel": {
        "template": load("templates/k8suniquelabel_template.yaml")[0],
        "constraint": load("constraints/all_cm_gatekeeper_label_unique.yaml")[0],
        "synced": load("good/no_dupe_cm.yaml"),
        "denied": load("bad/no_dupe_cm_2.yaml")[0],
        "source": "test/bats/test.bats:295-303 ('denied the request')",
    }
}
with open(os.path.join(HERE, "referential_vectors.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote referential_vectors.json")
