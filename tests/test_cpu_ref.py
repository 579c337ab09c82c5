"""The C++ CPU restatement that bench.py times as the reference arm (oracle/cpu_ref.cpp) against the Python oracle: same
violating pairs per constraint, same number of results, on the same objects."""
import json

from gatekeeper_b200 import workloads as W
from oracle import k8s
from oracle.cpu_ref import CpuRef


def _check(tm, cons, nss, blob, n, ep=k8s.AUDIT_EP):
    orc = k8s.Client()
    ref = CpuRef()
    for kind, rego, *rest in tm:
        orc.add_template(kind, rego, tuple(rest[0]) if rest and rest[0] else ())
        ref.add_template(kind, rego)
    for c in cons:
        orc.add_constraint(c)
        ref.add_constraint(c)
    for ns in nss:
        orc.add_namespace(ns)
        ref.add_namespace(ns)
    want_pairs, want_results, want_errs = {}, 0, 0
    for i in range(n):
        seen = set()
        for x in orc.review(k8s.Review(obj=json.loads(blob.get(i)), source="Original"), ep):
            key = "%s/%s" % x["constraint"]
            if x.get("autoreject"):
                want_errs += 1
                continue
            want_results += 1
            if key not in seen:
                seen.add(key)
                want_pairs[key] = want_pairs.get(key, 0) + 1
    pairs, nres, nerr, secs = ref.review_blob(blob, ep, threads=4)
    assert {k: v for k, v in pairs.items() if v} == want_pairs
    assert nres == want_results and nerr == want_errs
    ref.close()
    return nres


def test_cpu_ref_matches_oracle_config2():
    tm, cons = W.config2()
    n = 400
    assert _check(tm, cons, W.synth_namespaces(), W.synth_objects(0, n), n) > 2000


def test_cpu_ref_matches_oracle_without_namespace_cache():
    tm, cons = W.config2()
    n = 120
    _check(tm, cons, [], W.synth_objects(7000, n), n)


def test_cpu_ref_matches_oracle_config5_wildcards():
    tm, cons = W.config5()
    n = 300
    _check(tm, cons, W.synth_namespaces(), W.synth_objects(0, n, mode=W.MODE_CONFIG5 if hasattr(W, "MODE_CONFIG5") else 0), n)
