#!/usr/bin/env python
"""Generates tests/golden/gator_test_table.json from pkg/gator/test/test_test.go:85-268 (TestTest): for each row the
input documents (the YAML constants of pkg/gator/fixtures/fixtures.go the row names) and the exact results the reference
asserts: message, constraint name, enforcement action and scoped actions where the row states them.  Data only.

Run in the authoring container only (needs /root/reference):  python tests/golden/make_gator_table.py
"""
import json
import os
import re
import sys

import yaml

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_match_vectors import REF, val  # noqa: E402
from make_target_vectors import table  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    fx = {}
    src = open(f"{REF}/pkg/gator/fixtures/fixtures.go").read()
    for m in re.finditer(r"\t(\w+)\s*=\s*`\n?(.*?)`", src, re.S):
        try:
            doc = yaml.safe_load(m.group(2))
        except yaml.YAMLError:
            continue
        if isinstance(doc, dict):
            fx[m.group(1)] = doc
    tsrc = open(f"{REF}/pkg/gator/test/test_test.go").read()
    # `constraintNeverValidate, err = reader.ReadUnstructured([]byte(fixtures.ConstraintNeverValidate))` -- init(), :34-68
    var_of = dict(re.findall(r"(\w+), err = reader\.ReadUnstructured\(\[\]byte\(fixtures\.(\w+)\)\)", tsrc))
    rows, line = table(tsrc, "TestTest", "tcs := []struct")
    out = []
    for _, row in rows:
        f = {a[1]: b for a, b in row[2]}
        inputs = [b[1].split(".")[-1] for _, b in f["inputs"][2]]
        want = []
        for _, g in (f["want"][2] if "want" in f else []):
            res = {a[1]: b for a, b in {a[1]: b for a, b in g[2]}["Result"][2]}
            w = {"msg": val(res["Msg"]), "constraint": fx[var_of[res["Constraint"][1]]]["metadata"]["name"],
                 "constraintKind": fx[var_of[res["Constraint"][1]]]["kind"]}
            if "EnforcementAction" in res:
                w["enforcementAction"] = val(res["EnforcementAction"])
            if "ScopedEnforcementActions" in res:
                w["scopedEnforcementActions"] = val(res["ScopedEnforcementActions"])
            want.append(w)
        err = f.get("err")
        out.append({"source_test": "pkg/gator/test/test_test.go:TestTest (table at line %d)" % line, "name": val(f["name"]), "inputs": inputs,
                    "docs": [fx[i] for i in inputs], "want": want, "wantErr": None if err is None else err[1]})
    json.dump(out, open(os.path.join(HERE, "gator_test_table.json"), "w"), indent=1, sort_keys=True, default=str)
    print(len(out), "rows")
    for r in out:
        print(" ", r["name"], r["inputs"], len(r["want"]), r["wantErr"])


if __name__ == "__main__":
    main()
