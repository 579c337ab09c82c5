"""SURVEY 8 rows a-3 (audit aggregation), a-4 (excluder), a-14 (admission messages), a-17 (bench metric definitions):
the engine's host code against the oracle, on the test-only CPU backend.  The same cases run on the GPU in test_gpu.py."""
import json

import parity_cases as P
from conftest import HOSTEMU

LIB = HOSTEMU   # tests/test_audit_gpu.py re-runs this module against the CUDA library (LIB = None)
from gatekeeper_b200 import metrics as M
from oracle import audit as OA

MS = 1_000_000


def test_audit_aggregation_matches_oracle():
    P.case_audit(LIB, n=1200)


def test_audit_small_limit():
    P.case_audit(LIB, n=400, limit=2, excluded=())


def test_validation_messages():
    P.case_validation_messages(LIB)


def test_truncate_string_reference_behaviour():
    # pkg/audit/manager.go:1043-1052: > size => first size-3 bytes + "..."; size <= 3 keeps `size` bytes + "..."
    assert OA.truncate_string("a" * 256) == "a" * 256
    assert OA.truncate_string("a" * 257) == "a" * 253 + "..."
    assert OA.truncate_string("abcdef", 3) == "abc..."


def test_percentile_vectors_of_the_reference():
    # pkg/gator/bench/metrics_test.go:75-147 (tolerance there: 1 ms)
    five = [10 * MS, 20 * MS, 30 * MS, 40 * MS, 50 * MS]
    assert M.percentile([], 50) == 0
    assert M.percentile([100 * MS], 50) == 100 * MS
    assert M.percentile(five, 50) == 30 * MS
    assert abs(M.percentile(five, 99) - 49600 * 1000) <= MS
    assert M.percentile([10 * MS, 20 * MS, 30 * MS], 100) == 30 * MS
    assert M.percentile([10 * MS, 20 * MS], 0) == 10 * MS


def test_latencies_and_throughput_vectors_of_the_reference():
    # pkg/gator/bench/metrics_test.go:8-73,151-176
    assert M.calculate_latencies([]) == {"min": 0, "max": 0, "mean": 0, "p50": 0, "p95": 0, "p99": 0}
    one = M.calculate_latencies([100 * MS])
    assert (one["min"], one["max"], one["mean"]) == (100 * MS,) * 3
    for ds in ([10, 20, 30, 40, 50], [50, 10, 30, 20, 40]):
        l = M.calculate_latencies([d * MS for d in ds])
        assert (l["min"], l["max"], l["mean"]) == (10 * MS, 50 * MS, 30 * MS)
    assert M.calculate_throughput(100, 0) == 0
    assert M.calculate_throughput(100, 1000 * MS) == 100
    assert M.calculate_throughput(50, 500 * MS) == 100


def test_excluder_vectors_of_the_reference():
    """pkg/controller/config/process/excluder_test.go:11-66 through the oracle and through the engine's excluder stage."""
    from conftest import golden, make_pair
    from gatekeeper_b200 import driver as D
    from oracle import k8s
    t = golden("templates.json")["fixtures_TemplateNeverValidate"]
    for v in golden("target_vectors.json")["excluder"]:
        pod = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "p", "namespace": v["namespace"]}}
        nsobj = {"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": v["namespace"]}}
        assert k8s.is_namespace_excluded(v["patterns"], pod) == v["excluded"], v["name"]
        assert k8s.is_namespace_excluded(v["patterns"], nsobj) == v["excluded"], v["name"]      # a Namespace is tested by its own name
        orc, drv, _ = make_pair([(t["kind"], t["rego"])], [{"kind": t["kind"], "metadata": {"name": "c"}}], lib_path=LIB)
        drv.SetExcludedNamespaces("audit", v["patterns"])
        for o in (pod, nsobj):
            resp = drv.ReviewBatch([D.Review(object=o, source="Original")], k8s.AUDIT_EP, process="audit")
            assert (len(resp.results) == 0) == v["excluded"], (v["name"], o["kind"])
            resp = drv.ReviewBatch([D.Review(object=o, source="Original")], k8s.AUDIT_EP, process="webhook")   # other process: not excluded
            assert len(resp.results) == 1


def test_audit_from_cache_scenarios_of_the_reference():
    """pkg/audit/manager_test.go:103-171 Test_auditFromCache (transcribed; fakes from pkg/fakes/fixtures.go:14-90): one Pod in
    test-namespace-1, the deny-all template; violations: excluded namespace 0, not excluded 1, constraint scoped to the
    webhook point only 0, scoped to the audit point 1.  Through the oracle and through the engine's audit aggregation."""
    from conftest import make_pair
    from gatekeeper_b200 import driver as D
    from oracle import audit as OA
    from oracle import k8s
    rego = 'package goodrego\n\nviolation[{"msg": msg}] {\n   msg := "denyall"\n}'
    pod = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "test-pod", "namespace": "test-namespace-1"}}
    def scoped(ep):
        return {"apiVersion": "constraints.gatekeeper.sh/v1beta1", "kind": "denyall", "metadata": {"name": "constraint"},
                "spec": {"enforcementAction": "scoped", "scopedEnforcementActions": [
                    {"enforcementPoints": [{"name": ep}], "action": "deny"}, {"enforcementPoints": [{"name": ep}], "action": "warn"}]}}
    denyall = {"apiVersion": "constraints.gatekeeper.sh/v1beta1", "kind": "denyall", "metadata": {"name": "constraint"}}
    for name, excluded, con, want in (("obj excluded from audit", ["test-namespace-1"], denyall, 0), ("obj not excluded from audit", [], denyall, 1),
                                      ("audit excluded from constraint", [], scoped(k8s.WEBHOOK_EP), 0),
                                      ("audit included in constraints", [], scoped(k8s.AUDIT_EP), 1)):
        orc, drv, _ = make_pair([("denyall", rego)], [con], lib_path=LIB)
        assert len(OA.audit(orc, [pod], excluded_namespaces=excluded)["results"]) == want, name
        drv.SetExcludedNamespaces("audit", excluded)
        run = D.AuditRun(drv)
        run.add_batch(drv.upload([D.Review(object=pod, source="Original")], process="audit"), k8s.AUDIT_EP)
        rep = run.report()
        assert rep["results"] == want, (name, rep)
        if want and con is not denyall:
            v = rep["violations"]["denyall/constraint"][0]
            assert v["enforcementAction"] == "scoped" and v["enforcementActions"] == ["deny", "warn"]


def test_limit_queue_vectors_of_the_reference_through_the_engine():
    """pkg/audit/manager_test.go:40-103 (Test_newSVQueue / Test_LimitQueue): three violations on objects of the three GVKs the
    reference uses; limit 3 pops sv3, sv1, sv2 -- limit 2 keeps and pops sv1, sv2.  Here as the order of the status list."""
    from conftest import make_pair
    from gatekeeper_b200 import driver as D
    from oracle import k8s
    rego = 'package p\nviolation[{"msg": "m"}] { true }'
    objs = [{"apiVersion": "rbac.authorization.k8s.io/v1", "kind": "ClusterRoleBinding", "metadata": {"name": "x"}},
            {"apiVersion": "authorization.k8s.io/v1", "kind": "SubjectAccessReview", "metadata": {"name": "x"}},
            {"apiVersion": "rbac.authorization.k8s.io/v1", "kind": "RoleBinding", "metadata": {"name": "x"}}]
    orc, drv, _ = make_pair([("P", rego)], [{"kind": "P", "metadata": {"name": "c"}}], lib_path=LIB)
    for limit, want in ((3, ["RoleBinding", "ClusterRoleBinding", "SubjectAccessReview"]), (2, ["ClusterRoleBinding", "SubjectAccessReview"])):
        run = D.AuditRun(drv, violations_limit=limit)
        run.add_batch(drv.upload([D.Review(object=o, source="Original") for o in objs]), k8s.AUDIT_EP)
        rep = run.report()
        assert [v["kind"] for v in rep["violations"]["P/c"]] == want
        assert rep["totalViolations"] == {"P/c": 3}


GVM_VECTORS = [   # pkg/webhook/policy_test.go:840-948 TestGetValidationMessages: (name, result actions, deny count, warn count)
    ("Only One Dry Run", ["dryrun"], 0, 0), ("Only One Deny", ["deny"], 1, 0), ("Only One Warn", ["warn"], 0, 1),
    ("One Dry Run and One Deny", ["dryrun", "deny"], 1, 0), ("One Dry Run, One Deny, One Warn", ["dryrun", "deny", "warn"], 1, 1),
    ("Two Deny", ["deny", "deny"], 2, 0), ("Two Warn", ["warn", "warn"], 0, 2), ("Two Dry Run", ["dryrun", "dryrun"], 0, 0),
    ("Random EnforcementAction", ["random"], 0, 0)]


def test_get_validation_messages_vectors_of_the_reference():
    """The nine action mixes the reference pins (message counts), through the oracle's restatement on bare results and
    through the engine: one deny-all constraint per listed action, the reference's nameless Namespace request."""
    from conftest import make_pair
    from gatekeeper_b200 import driver as D
    from oracle import k8s
    rego = 'package foo\nviolation[{"msg": "test"}] { true }'
    ns = {"apiVersion": "v1", "kind": "Namespace"}
    for name, actions, n_deny, n_warn in GVM_VECTORS:
        bare = [{"msg": "test", "constraint": ("Foo", "ph"), "enforcementAction": a if a != "random" else "unrecognized", "scopedEnforcementActions": []}
                for a in actions]
        d, w = k8s.validation_messages(bare)
        assert (len(d), len(w)) == (n_deny, n_warn), name
        assert all(m == "[ph] test" for m in d + w)
        cons = [{"kind": "Foo", "metadata": {"name": "ph-%d" % i}, "spec": {"enforcementAction": a}} for i, a in enumerate(actions)]
        orc, drv, _ = make_pair([("Foo", rego)], cons, lib_path=LIB)
        (deny, warn), = drv.ValidationMessages([D.Review(object=ns, namespace_name="")])
        assert (len(deny), len(warn)) == (n_deny, n_warn), (name, deny, warn)
        assert all(m.startswith("[ph-") and m.endswith("] test") for m in deny + warn)


def test_webhook_excluded_namespaces_vectors_of_the_reference():
    """pkg/webhook/policy_test.go:422-523 TestExcludedNamespaces (transcribed): Config excludes "kube-*" for every process; the
    deny-everything template K8sGoodRego (:157-200); the request's Namespace decides (the object's own is empty,
    pkg/webhook/common.go:181); DELETE is judged on OldObject and is an error without one."""
    from conftest import make_pair
    from gatekeeper_b200 import driver as D
    from oracle import k8s
    rego = 'package goodrego\n\nviolation[{"msg": msg}] {\n   msg := "Maybe this will work?"\n}'
    raw = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "acbd", "namespace": ""}}
    orc, drv, _ = make_pair([("K8sGoodRego", rego)], [{"kind": "K8sGoodRego", "metadata": {"name": "constraint"}}], lib_path=LIB)
    drv.SetExcludedNamespaces("*", ["kube-*"])
    for name, ns, op, obj, old, allowed in (("ExcludedNamespace invalid create", "notkube-test", "CREATE", raw, None, False),
                                            ("ExcludedNamespace valid create", "kube-test", "CREATE", raw, None, True),
                                            ("ExcludedNamespace invalid delete", "kube-test", "DELETE", None, None, False),
                                            ("ExcludedNamespace valid delete", "kube-test", "DELETE", None, raw, True)):
        rev = D.Review(object=obj, old_object=old, operation=op, namespace_name=ns)
        resp = drv.ReviewBatch([rev], k8s.WEBHOOK_EP, process="webhook")
        err = (resp.object_errors or [None])[0]
        denied = bool(err) or any(r.enforcement_action == "deny" for r in resp.results)
        assert (not denied) == allowed, (name, err, [r.msg for r in resp.results])
        if name.endswith("invalid delete"):
            assert err and "oldObject" in err          # errOldObjectIsNil
        # the oracle: the excluder stage, then Client.Review
        target_obj = old if op == "DELETE" else obj
        if target_obj is None:
            continue
        probe = dict(target_obj, metadata=dict(target_obj["metadata"], namespace=ns))
        o_allowed = k8s.is_namespace_excluded(["kube-*"], probe) or not orc.review(
            k8s.Review(obj=obj, old=old, operation=op, namespace=ns), k8s.WEBHOOK_EP)
        assert o_allowed == allowed, name


def test_admission_coalescer_batches_concurrent_reviews():
    """gk_coalescer_*: 48 threads each send one admission review; every caller must get exactly its own request's outcome
    (compared with a direct ReviewBatch of the same request), and the requests must have been gathered into micro-batches."""
    import threading
    from conftest import golden, make_pair
    from gatekeeper_b200 import driver as D
    from oracle import k8s
    psp = golden("psp_suite.json")
    cons = [json.loads(json.dumps(c)) for c in psp["constraints"]]
    cons[0].setdefault("spec", {})["enforcementAction"] = "warn"
    orc, drv, _ = make_pair([(t["kind"], t["rego"]) for t in psp["templates"]], cons, lib_path=LIB)
    pods = psp["pods"]
    ok_pod = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "fine", "namespace": "default"}, "spec": {"containers": [{"name": "c", "image": "x"}]}}
    reqs = [D.Review(object=(ok_pod if i % 7 == 0 else pods[i % len(pods)]), operation="CREATE", namespace_name="default") for i in range(48)]
    co = D.Coalescer(drv, max_batch=16, max_wait_us=20000)
    out = [None] * len(reqs)
    def worker(i):
        out[i] = co.review(reqs[i])
    th = [threading.Thread(target=worker, args=(i,)) for i in range(len(reqs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    st = co.stats()
    assert st["reviews"] == len(reqs) and st["batches"] < len(reqs), st          # some coalescing happened
    assert max(o["batch_size"] for o in out) > 1 and max(o["batch_size"] for o in out) <= 16
    for i, r in enumerate(reqs):
        direct = drv.ReviewBatch([r], k8s.WEBHOOK_EP, process="webhook")
        want = sorted((x.constraint, x.msg, x.enforcement_action) for x in direct.results)
        got = sorted((x["constraint"], x["msg"], x["enforcementAction"]) for x in out[i]["results"])
        assert got == want, i
        (deny, warn), = drv.ValidationMessages([r])
        assert sorted(out[i]["messages"]["deny"]) == sorted(deny) and sorted(out[i]["messages"]["warn"]) == sorted(warn)
        assert out[i]["error"] is None
    assert any(o["messages"]["deny"] for o in out) and any(not o["results"] for o in out)
    co.close()
