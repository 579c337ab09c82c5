"""The oracle pinned against the reference's own golden vectors.  CPU only."""
import json

import pytest

from conftest import golden
from oracle import audit, k8s, rego


# ---- Matcher: fully pinned by in-tree tests -------------------------------------------------------------
@pytest.mark.parametrize("v", golden("match_vectors.json"), ids=lambda v: v["name"][:60])
def test_match_vectors(v):
    """pkg/mutation/match/match_test.go:17-683 (TestMatch) and :847-1040 (Test_namesMatch)."""
    try:
        got, err = k8s.matches(v["match"], v["object"], v["namespace"], v["source"]), False
    except k8s.MatchError:
        got, err = False, True
    assert (got, err) == (v["wantMatch"], v["wantErr"])


@pytest.mark.parametrize("w", golden("wildcard_vectors.json"), ids=lambda w: f"{w['fn']}:{w['name']}"[:60])
def test_wildcard_vectors(w):
    """pkg/wildcard/wildcard_test.go:7-193."""
    f = k8s.wildcard_matches if w["fn"] == "matches" else k8s.wildcard_matches_generate_name
    assert f(w["w"], w["candidate"]) == w["matches"]


def test_matcher_old_or_new_object():
    """Matcher.Match is true if EITHER object matches; error only if both are nil -- pkg/target/matcher.go:44-71,
    vectors in pkg/target/target_test.go:657-980."""
    m = {"namespaces": ["a"]}
    mk = lambda ns: {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "p", "namespace": ns}}
    assert k8s.matcher_match(m, k8s.Review(obj=mk("a"), old=mk("b")))
    assert k8s.matcher_match(m, k8s.Review(obj=mk("b"), old=mk("a")))
    assert not k8s.matcher_match(m, k8s.Review(obj=mk("b"), old=mk("c")))
    with pytest.raises(k8s.MatchError, match="neither object nor old object"):
        k8s.matcher_match(m, k8s.Review(obj=None, old=None))
    assert k8s.matcher_match(None, k8s.Review(obj=None, old=None))   # nil match: no-op (matcher.go:22-25)


def test_excluder():
    """pkg/controller/config/process/excluder.go:95-127."""
    pod = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "p", "namespace": "kube-system"}}
    ns = {"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": "kube-public"}}
    assert k8s.is_namespace_excluded(["kube-*"], pod)
    assert k8s.is_namespace_excluded(["kube-*"], ns)
    assert not k8s.is_namespace_excluded(["gatekeeper-*"], pod)


# ---- Rego results pinned by the reference's golden messages -------------------------------------------------
def _run_docs(docs, ep=k8s.GATOR_EP):
    c = k8s.Client()
    for d in docs:
        if d.get("kind") == "ConstraintTemplate":
            c.add_template(*k8s.template_from_yaml_obj(d))
    for d in docs:
        if str(d.get("apiVersion", "")).startswith("constraints.gatekeeper.sh"):
            c.add_constraint(d)
    for d in docs:
        if d.get("kind") == "Namespace":
            c.add_namespace(d)
    out = []
    for d in docs:   # gator test reviews EVERY input object (pkg/gator/test/test.go:109-173)
        out += c.review(k8s.Review(obj=d), ep)
    return out


@pytest.mark.parametrize("case", golden("gator_cases.json"), ids=lambda c: c["name"])
def test_gator_golden_messages(case):
    """test/gator/test/test.bats:73,95,114,165-202 -- exact violation messages."""
    res = _run_docs(case["docs"])
    msgs = [r["msg"] for r in res]
    if case["must_contain"] is None:
        assert msgs, "rego v1 template manifest must produce violations (test.bats:88-90)"
        return
    for m in case["must_contain"]:
        assert m in msgs
    if not case["must_contain"]:
        assert not [r for r in res if r["enforcementAction"] == "deny"]


def test_required_labels_message_format():
    """`you must provide labels: {"geo"}` (test/gator/test/test.bats:233,249) and
    'you must provide labels: {"gatekeeper"}' (website/docs/audit.md:52): %v of a set of strings."""
    src = golden("templates.json")["requiredlabels_regov1"]["rego"]
    m = rego.Module(src)
    inp = rego.from_json({"review": {"object": {"metadata": {"labels": {"a": "b"}}}}, "parameters": {"labels": ["geo"]}})
    vs = rego.eval_violations(m, inp)
    assert [v["msg"] for v in vs] == ['you must provide labels: {"geo"}']
    assert rego.to_json(vs[0]["details"]) == {"missing_labels": ["geo"]}
    assert rego.fmt_value(frozenset(["b", "a"]), top=True) == '{"a", "b"}'
    assert rego.fmt_value(frozenset(), top=True) == "set()"


def test_psp_suite_each_pod_violates_its_template():
    """pkg/webhook/policy_benchmark_test.go:264-271: 'all constraints applicable and all requests violating'."""
    psp = golden("psp_suite.json")
    c = k8s.Client()
    for t in psp["templates"]:
        c.add_template(t["kind"], t["rego"])
    for x in psp["constraints"]:
        c.add_constraint(x)
    want = {"nginx-host-filesystem": "K8sPSPHostFilesystem", "nginx-host-namespace": "K8sPSPHostNamespace",
            "nginx-host-networking-ports": "K8sPSPHostNetworkingPorts", "nginx-privileged": "K8sPSPPrivilegedContainer",
            "nginx-volume-types": "K8sPSPVolumeTypes"}
    for pod in psp["pods"]:
        kinds = {r["constraint"][0] for r in c.review(k8s.Review(obj=pod), k8s.WEBHOOK_EP)}
        assert want[pod["metadata"]["name"]] in kinds


def test_verify_suite_fooisbar():
    """test/gator/verify/suite.yaml:1-37: allow/deny counts incl. scoped enforcement points."""
    vs = golden("verify_suite.json")
    files = vs["files"]
    tmpl = k8s.template_from_yaml_obj(vs["template"])

    def run(constraint_file, obj_file):
        c = k8s.Client()
        c.add_template(*tmpl)
        c.add_constraint(files[constraint_file][0])
        return c.review(k8s.Review(obj=files[obj_file][0]), k8s.GATOR_EP)

    assert run("constraint.yaml", "allow_foo.yaml") == []
    assert len(run("constraint.yaml", "deny_foo.yaml")) >= 1
    assert run("constraint_with_scopedEA.yaml", "allow_foo.yaml") == []
    assert len(run("constraint_with_scopedEA.yaml", "deny_foo.yaml")) >= 1
    assert run("constraint_with_scopedEA_without_gator_ep.yaml", "deny_foo.yaml") == []


def test_template_libs_container_limits():
    """test/bats/test.bats:268-279 with test/bats/tests/templates/k8scontainterlimits_template.yaml (`libs`: package
    lib.helpers imported as data.lib.helpers): bad/opa_no_limits.yaml is denied, good/opa.yaml is admitted; plus the helper
    functions' arithmetic on hand-checked quantities."""
    g = golden("libs_template.json")
    tmpl = k8s.template_from_yaml_obj(g["template"])
    assert len(tmpl) == 3 and tmpl[2][0].lstrip().startswith("package lib.helpers")
    c = k8s.Client()
    c.add_template(*tmpl)
    c.add_constraint(g["constraint"])      # cpu 200m, memory 1Gi
    assert [r["msg"] for r in c.review(k8s.Review(obj=g["denied"]), k8s.WEBHOOK_EP)] == ["container <opa> has no resource limits"]
    assert c.review(k8s.Review(obj=g["admitted"]), k8s.WEBHOOK_EP) == []

    def msgs(cpu, mem):
        pod = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "p", "namespace": "gatekeeper-test-playground"},
               "spec": {"containers": [{"name": "c", "resources": {"limits": {"cpu": cpu, "memory": mem}}}]}}
        return sorted(r["msg"] for r in c.review(k8s.Review(obj=pod), k8s.WEBHOOK_EP))

    assert msgs("200m", "1Gi") == []
    assert msgs("201m", "1Gi") == ["container <c> cpu limit <201m> is higher than the maximum allowed of <200m>"]
    assert msgs("0.1", "1Gi") == ["container <c> cpu limit <0.1> could not be parsed"]          # canonify_cpu: digits only, or <n>m
    assert msgs(1, "1024Mi") == ["container <c> cpu limit <1> is higher than the maximum allowed of <200m>"]
    assert msgs("100m", "1025Mi") == ["container <c> memory limit <1025Mi> is higher than the maximum allowed of <1Gi>"]
    assert msgs("100m", "1G") == []                                                              # 10^12 < 2^30 * 1000
    assert msgs("100m", "") == ["container <c> has no memory limit", "container <c> memory limit <> could not be parsed"]
    # a lib must live under `package lib...`; an import must name a lib of the template
    with pytest.raises(rego.RegoError, match="rego_compile_error"):
        k8s.Client().add_template("X", tmpl[1], ["package helpers\nf(x) = x { true }\n"])
    with pytest.raises(rego.RegoError, match="rego_compile_error"):
        k8s.Client().add_template("X", tmpl[1], [])


# pkg/util/enforcement_action_test.go:113-165 (TestGetEnforcementAction) and :235-385 (TestScopedActionForEP), hand-translated
EA_VECTORS = [({}, "deny"), ({"spec": {"enforcementAction": "notsupported"}}, "unrecognized"), ({"spec": {"enforcementAction": "dryrun"}}, "dryrun")]
AUDIT, WEBHOOK, ALL = "audit.gatekeeper.sh", "validation.gatekeeper.sh", "*"


def _sea(*pairs):
    return {"spec": {"scopedEnforcementActions": [{"action": a, "enforcementPoints": [{"name": n} for n in eps]} for a, eps in pairs]}}


SCOPED_VECTORS = [
    ("valid enforcement point", AUDIT, _sea(("deny", [AUDIT]), ("warn", [WEBHOOK])), ["deny"]),
    ("multiple enforcement points", WEBHOOK, _sea(("deny", [AUDIT, WEBHOOK]), ("warn", [WEBHOOK])), ["deny", "warn"]),
    ("no matching enforcement point", AUDIT, _sea(("deny", [WEBHOOK]), ("warn", [WEBHOOK])), []),
    ("wildcard enforcement point", AUDIT, _sea(("deny", [ALL]), ("warn", [WEBHOOK])), ["deny"]),
    ("missing scopedEnforcementActions", AUDIT, {"spec": {}}, []),
]


def test_enforcement_action_vectors():
    for item, want in EA_VECTORS:
        assert k8s.get_enforcement_action(item) == want
    for name, ep, item, want in SCOPED_VECTORS:
        assert k8s.scoped_actions_for_ep(ep, item) == want, name


def test_validate_constraint_vectors():
    """pkg/target/target_test.go:42-399 (TestValidateConstraint): the 11 cases, error expected or not."""
    g = golden("validate_constraint_vectors.json")
    assert len(g["cases"]) == 11
    for cse in g["cases"]:
        if cse["error_expected"]:
            with pytest.raises(k8s.ValidateError):
                k8s.validate_constraint(cse["constraint"])
        else:
            k8s.validate_constraint(cse["constraint"])


def test_fixture_templates():
    """pkg/gator/test/test_test.go:85-452: "never validate" x N, first/second message, compile error."""
    t = golden("templates.json")
    obj = {"apiVersion": "v1", "kind": "Object", "metadata": {"name": "o"}}

    def msgs(name, params=None, review=None):
        m = rego.Module(t[name]["rego"])
        inp = rego.from_json({"review": review or {"object": obj, "kind": {"kind": "Object"}}, "parameters": params or {}})
        return sorted(v["msg"] for v in rego.eval_violations(m, inp))

    assert msgs("fixtures_TemplateNeverValidate") == ["never validate"]
    assert msgs("fixtures_TemplateAlwaysValidate") == []
    assert msgs("fixtures_TemplateNeverValidateTwice") == ["first message", "second message"]
    with pytest.raises(rego.RegoError, match="unsafe"):
        rego.Module(t["fixtures_TemplateCompileError"]["rego"])
    assert msgs("fixtures_TemplateValidateUserInfo", review={"object": obj, "userInfo": {"username": "bob"}}) == \
        ["username is not allowed to perform this operation: bob"]
    assert msgs("fixtures_TemplateValidateUserInfo", review={"object": obj, "userInfo": {"username": "system:foo"}}) == []


def test_namespace_selector_missing_namespace_is_a_violation():
    """pkg/gator/verify/runner_test.go:986-989: matcher error => `Violations: yes`, message has "missing Namespace"."""
    t = golden("templates.json")["fixtures_TemplateNeverValidate"]
    c = k8s.Client()
    c.add_template(t["kind"], t["rego"])
    c.add_constraint({"kind": t["kind"], "metadata": {"name": "c"},
                      "spec": {"match": {"namespaceSelector": {"matchLabels": {"bar": "qux"}}}}})
    res = c.review(k8s.Review(obj={"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "p", "namespace": "ns1"}}), k8s.GATOR_EP)
    assert len(res) == 1 and "missing Namespace" in res[0]["msg"]


# ---- audit aggregation ---------------------------------------------------------------------------------
def test_limit_queue_matches_reference_test():
    """pkg/audit/manager_test.go:39-101 (Test_SVQueue, Test_LimitQueue) restated."""
    mk = lambda g, k: {"group": g, "version": "v1", "kind": k, "namespace": "", "name": "", "message": "", "enforcementAction": ""}
    sv1, sv2, sv3 = mk("rbac.authorization.k8s.io", "ClusterRoleBinding"), mk("authorization.k8s.io", "SubjectAccessReview"), \
        mk("rbac.authorization.k8s.io", "RoleBinding")
    q = audit.LimitQueue(3)
    for s in (sv1, sv2, sv3):
        q.push(s)
    assert q.drain_descending() == [sv3, sv1, sv2]
    q = audit.LimitQueue(2)
    for s in (sv1, sv2, sv3):
        q.push(s)
    assert q.drain_descending() == [sv1, sv2]


def test_truncate_string():
    """pkg/audit/manager.go:1043-1052, test at pkg/audit/manager_test.go:229."""
    assert audit.truncate_string("a" * 300) == "a" * 253 + "..."
    assert audit.truncate_string("short") == "short"
    assert audit.truncate_string("abcdef", 3) == "abc..."


def test_rego_semantics_risk_list():
    """undefined vs false, set semantics, multi-body OR, number formatting."""
    m = rego.Module("""package t
f(x) = 1 { x == "a" }
f(x) = 2 { x == "b" }
violation[{"msg": msg}] { not input.review.object.missing; msg := "undefined is not-able" }
violation[{"msg": msg}] { input.review.object.zero; msg := "zero is truthy" }
violation[{"msg": msg}] { input.review.object.f; msg := "false is truthy" }
violation[{"msg": msg}] { x := input.review.object.list[_]; msg := sprintf("dup %v", [x]) }
violation[{"msg": msg}] { msg := sprintf("%v %v %v", [f("b"), 10 * 100, {"k": [1, "s"]}]) }
""")
    inp = rego.from_json({"review": {"object": {"zero": 0, "f": False, "list": [1, 1, 1.0]}}, "parameters": {}})
    msgs = sorted(v["msg"] for v in rego.eval_violations(m, inp))
    assert msgs == ['2 1000 {"k": [1, "s"]}', "dup 1", "undefined is not-able", "zero is truthy"]


# ---- pkg/target: the reference's own Review-boundary scenarios and Matcher.Match vectors (tests/golden/make_target_vectors.py)
def test_target_constraint_enforcement_scenarios():
    """pkg/target/target_integration_test.go:164-520 -- 26 scenarios x 3 request shapes through the real client + Rego
    driver in the reference: `allowed` <=> Review returns no results."""
    from parity_cases import DENY_ALL, _target_shapes
    n = 0
    for v in golden("target_vectors.json")["constraint_enforcement"]:
        c = k8s.Client()
        c.add_template("DenyAll", DENY_ALL)
        c.add_constraint(v["constraint"])
        for shape, kw in _target_shapes(v):
            rv = k8s.Review(obj=kw.get("object"), old=kw.get("old_object"), ns=kw.get("namespace"), namespace=kw.get("namespace_name"))
            res = c.review(rv, k8s.AUDIT_EP)
            assert (len(res) == 0) == v["allowed"], (v["name"], shape, res)
            n += 1
    assert n == 78


def test_target_matcher_match_vectors():
    """pkg/target/target_test.go:657-914 TestMatcher_Match."""
    n = 0
    for v in golden("target_vectors.json")["matcher_match"]:
        req = v["request"]
        if req is None or v["match"] is None:
            continue
        cache = {v["cachedNamespace"]["metadata"]["name"]: v["cachedNamespace"]} if v["cachedNamespace"] else {}
        obj = req.get("object")
        if v["wantErr"] == "ErrRequestObject" and obj is not None and "kind" not in obj:
            continue    # "Raw object doesn't unmarshal": rejected while decoding the request, before Matcher.Match
        nsn = req.get("namespaceName", "") if req["shape"] in ("AugmentedReview", "AdmissionRequest") else None
        rv = k8s.Review(obj=obj, old=req.get("oldObject"), ns=req.get("namespace"), namespace=nsn)
        try:
            got, err = k8s.matcher_match(v["match"], rv, cache), None
        except k8s.MatchError as e:
            got, err = False, str(e)
        if v["wantErr"] == "ErrMatching":
            assert err and err.startswith("error matching the requested object"), v["name"]
        elif v["wantErr"] == "ErrRequestObject":
            assert err and err.startswith("invalid request object"), v["name"]
        else:
            assert err is None and got == v["want"], (v["name"], got, err)
        n += 1
    assert n >= 14


def test_process_data_paths():
    """pkg/target/target_test.go:401-495 TestProcessData (the *unstructured.Unstructured rows; makeResource / makeNamespacedResource
    build {apiVersion, kind, metadata.name[, metadata.namespace]}): where a synced object lives under data.inventory -- the `path` of
    gk_add_data / Client.add_data (processUnstructured, pkg/target/target.go:40-66)."""
    res = lambda gv, kind, name, ns=None: {"apiVersion": gv, "kind": kind, "metadata": dict({"name": name}, **({"namespace": ns} if ns else {}))}
    assert k8s.Client.data_path(res("v1beta1", "Rock", "myrock")) == ["cluster", "v1beta1", "Rock", "myrock"]                      # "Cluster Object"
    assert k8s.Client.data_path(res("v1beta1", "Rock", "myrock", "foo")) == ["namespace", "foo", "v1beta1", "Rock", "myrock"]      # "Namespaced Object"
    assert k8s.Client.data_path(res("mygroup/v1beta1", "Rock", "myrock")) == ["cluster", "mygroup/v1beta1", "Rock", "myrock"]      # "Grouped Object"
    for bad, what in ((res("", "Rock", "myrock"), "has no version"), (res("v1beta1", "", "myrock"), "has no kind")):           # "No Version", "No Kind"
        with pytest.raises(ValueError) as e:
            k8s.Client.data_path(bad)
        assert "invalid request object" in str(e.value) and what in str(e.value)     # ErrRequestObject wrapped: target.go:43-48
