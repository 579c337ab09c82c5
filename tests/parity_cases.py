"""Parity cases shared by the CPU (host-emulation backend) and GPU (product library) test modules.
Every case loads the SAME templates / constraints / objects into the oracle and into the engine and
compares the full result set: (object, constraint, msg, details, enforcement action(s), autoreject)."""
import json
import os
import random

import numpy as np

from conftest import assert_same, engine_results, golden, make_pair, oracle_results
from gatekeeper_b200 import driver as D
from gatekeeper_b200 import workloads as W
from oracle import k8s


def _split_docs(docs):
    tm, cons, nss = [], [], []
    for d in docs:
        if d.get("kind") == "ConstraintTemplate":
            tm.append(k8s.template_from_yaml_obj(d))
        elif str(d.get("apiVersion", "")).startswith("constraints.gatekeeper.sh"):
            cons.append(d)
        elif d.get("kind") == "Namespace":
            nss.append(d)
    return tm, cons, nss


def case_gator(lib, case):
    tm, cons, nss = _split_docs(case["docs"])
    orc, drv, skipped = make_pair(tm, cons, nss, lib_path=lib, skip_unsupported=True)
    revs = [D.Review(object=d) for d in case["docs"]]
    resp = drv.ReviewBatch(revs, k8s.GATOR_EP)
    assert_same(oracle_results(orc, revs, k8s.GATOR_EP), engine_results(resp))
    if case["must_contain"] and not skipped:
        msgs = {r.msg for r in resp.results}
        for m in case["must_contain"]:
            assert m in msgs
    return resp


def case_psp(lib):
    psp = golden("psp_suite.json")
    orc, drv, _ = make_pair([(t["kind"], t["rego"]) for t in psp["templates"]], psp["constraints"], lib_path=lib)
    revs = [D.Review(object=p) for p in psp["pods"]]
    for ep in (k8s.WEBHOOK_EP, k8s.AUDIT_EP):
        resp = drv.ReviewBatch(revs, ep)
        want = oracle_results(orc, revs, ep)
        assert len(want) >= 5
        assert_same(want, engine_results(resp))
    return resp


def case_config2(lib, n, start=0, with_namespaces=True, ep=k8s.AUDIT_EP):
    tm, cons = W.config2()
    nss = W.synth_namespaces() if with_namespaces else []
    orc, drv, _ = make_pair(tm, cons, nss, lib_path=lib)
    blob = W.synth_objects(start, n)
    revs = [D.Review(object=json.loads(blob.get(i)), source="Original") for i in range(n)]
    resp = drv.ReviewBatch(revs, ep)
    want = oracle_results(orc, revs, ep)
    assert_same(want, engine_results(resp))
    # bitmap == set of violating (object, constraint) pairs; totals == column popcounts
    pairs = {(o, c) for (o, c, *_rest) in want if not _rest[-1]}
    assert resp.pairs() == pairs
    for ci, key in enumerate(resp.constraints):
        assert resp.totals[ci] == sum(1 for (_, c) in pairs if c == key)
    return resp, want


def case_mixed_kinds(lib, n):
    """config 4 shape: PSP suite over mixed GVKs (only Pods match `kinds`)."""
    tm, cons = W.config4()
    orc, drv, _ = make_pair(tm, cons, lib_path=lib)
    blob = W.synth_objects(0, n, mode=1)
    revs = [D.Review(object=json.loads(blob.get(i))) for i in range(n)]
    resp = drv.ReviewBatch(revs, k8s.AUDIT_EP)
    assert_same(oracle_results(orc, revs, k8s.AUDIT_EP), engine_results(resp))
    return resp


def case_config5(lib, n):
    tm, cons = W.config5()
    orc, drv, _ = make_pair(tm, cons, lib_path=lib)
    blob = W.synth_objects(1000, n)
    revs = [D.Review(object=json.loads(blob.get(i))) for i in range(n)]
    resp = drv.ReviewBatch(revs, k8s.AUDIT_EP)
    assert_same(oracle_results(orc, revs, k8s.AUDIT_EP), engine_results(resp))
    return resp


def case_allowedrepos_comprehension_variant(lib, n):
    """demo/agilebank variant: `satisfied := [good | repo = ...; good = startswith(...)]; not any(satisfied)`."""
    t = W.templates()["allowedrepos"]
    cons = [W._constraint(t["kind"], "repos", match=dict(W.POD), params={"repos": ["openpolicyagent/", "gcr.io/proj-0"]}),
            W._constraint(t["kind"], "none", params={"repos": []})]
    orc, drv, _ = make_pair([(t["kind"], t["rego"])], cons, lib_path=lib)
    blob = W.synth_objects(50, n)
    revs = [D.Review(object=json.loads(blob.get(i))) for i in range(n)]
    resp = drv.ReviewBatch(revs, k8s.AUDIT_EP)
    assert_same(oracle_results(orc, revs, k8s.AUDIT_EP), engine_results(resp))


def case_match_vectors(lib):
    """Every reference Matcher vector through the engine's in-kernel pre-filter (deny-all template): the
    object is flagged iff the reference says it matches; errors become autoreject results."""
    t = golden("templates.json")["fixtures_TemplateNeverValidate"]
    vectors = [v for v in golden("match_vectors.json") if v["object"] is not None]
    checked = 0
    for v in vectors:
        con = {"kind": t["kind"], "metadata": {"name": "c"}, "spec": {"match": v["match"]}}
        drv = D.Driver(lib_path=lib)
        drv.add_template(t["kind"], t["rego"])
        drv.AddConstraint(con)
        r = D.Review(object=v["object"], namespace=v["namespace"], source=v["source"])
        resp = drv.ReviewBatch([r], k8s.AUDIT_EP)
        flagged = bool(resp.viol_bits[0, 0] & 1)
        errored = bool(resp.err_bits[0, 0] & 1)
        assert (flagged, errored) == (v["wantMatch"], v["wantErr"]), v["name"]
        if errored:
            assert resp.results and resp.results[0].autoreject
        checked += 1
    assert checked >= 50


def case_admission_shapes(lib):
    """UPDATE with object+oldObject (either may match), DELETE (object := oldObject), explicit / cached /
    missing Namespace for namespaceSelector, user info -- pkg/target/target.go:86-138,262-280."""
    t = golden("templates.json")
    tm = [(t[n]["kind"], t[n]["rego"]) for n in ("fixtures_TemplateNeverValidate", "fixtures_TemplateValidateUserInfo", "namespacelabelcheck")]
    cons = [
        {"kind": "NeverValidate", "metadata": {"name": "only-a"}, "spec": {"match": {"namespaces": ["a"]}}},
        {"kind": "NeverValidate", "metadata": {"name": "nssel"}, "spec": {"match": {"namespaceSelector": {"matchLabels": {"bar": "qux"}}}}},
        {"kind": "NeverValidate", "metadata": {"name": "generated-only"}, "spec": {"match": {"source": "Generated"}}},
        {"kind": "NeverValidate", "metadata": {"name": "by-name"}, "spec": {"match": {"name": "web-*"}}},
        {"kind": "ValidateUserInfo", "metadata": {"name": "users"}, "spec": {"enforcementAction": "warn"}},
        {"kind": "K8sNamespaceLabelCheckRego", "metadata": {"name": "nslabel"}, "spec": {"parameters": {"requiredLabel": "bar"}}},
    ]
    ns_qux = {"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": "cached", "labels": {"bar": "qux"}}}
    orc, drv, _ = make_pair(tm, cons, [ns_qux], lib_path=lib)
    pod = lambda ns, name="p", gen=None: {"apiVersion": "v1", "kind": "Pod",
                                          "metadata": dict({"name": name, "namespace": ns}, **({"generateName": gen} if gen else {}))}
    revs = [
        D.Review(object=pod("a"), old_object=pod("b"), operation="UPDATE", source="Original"),
        D.Review(object=pod("b"), old_object=pod("a"), operation="UPDATE", source="Original"),
        D.Review(object=None, old_object=pod("a"), operation="DELETE", source="Original"),
        D.Review(object=pod("cached"), source="Generated"),
        D.Review(object=pod("uncached"), source="Original"),
        D.Review(object=pod("x"), namespace={"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": "x", "labels": {"bar": "qux"}}}, source="Original"),
        D.Review(object=pod("b", name="web-1"), source="Original", user_info={"username": "alice"}),
        D.Review(object=pod("b", name="", gen="web-"), source="Original", user_info={"username": "system:serviceaccount:x"}),
        D.Review(object={"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": "nsobj", "labels": {"bar": "qux"}}}, source="Original"),
        D.Review(object=pod("b"), source=""),   # source unset vs matcher Generated => error (match.go:244-246)
    ]
    for ep in (k8s.WEBHOOK_EP, k8s.AUDIT_EP):
        resp = drv.ReviewBatch(revs, ep)
        assert_same(oracle_results(orc, revs, ep), engine_results(resp))
    return resp


def case_review_errors(lib):
    """Bad JSON / missing kind / DELETE without oldObject are per-object review errors, not violations
    (pkg/target/matcher.go:73-93, pkg/target/target.go:262-270); the rest of the batch is unaffected."""
    t = golden("templates.json")["fixtures_TemplateNeverValidate"]
    drv = D.Driver(lib_path=lib)
    drv.add_template(t["kind"], t["rego"])
    drv.AddConstraint({"kind": t["kind"], "metadata": {"name": "c"}, "spec": {}})
    ok = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "p"}}
    revs = [D.Review(object=ok), D.Review(object=b"{not json"), D.Review(object={"apiVersion": "v1", "metadata": {}}),
            D.Review(object=None, old_object=None, operation="DELETE"), D.Review(object=ok)]
    resp = drv.ReviewBatch(revs, k8s.AUDIT_EP)
    assert [bool(e) for e in resp.object_errors] == [False, True, True, True, False]
    assert sorted(r.object for r in resp.results) == [0, 4]


def case_edge_batches(lib):
    """Empty batch, single object, no constraints, >32 constraints (two bitmap words), constant predicates."""
    t = golden("templates.json")
    drv = D.Driver(lib_path=lib)
    ok = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "p"}}
    resp = drv.ReviewBatch([D.Review(object=ok)], k8s.AUDIT_EP)      # no templates at all
    assert resp.results == [] and resp.totals == []
    drv.add_template(t["fixtures_TemplateNeverValidate"]["kind"], t["fixtures_TemplateNeverValidate"]["rego"])
    drv.add_template(t["fixtures_TemplateAlwaysValidate"]["kind"], t["fixtures_TemplateAlwaysValidate"]["rego"])
    for i in range(40):
        kind = "NeverValidate" if i % 3 else "AlwaysValidate"
        drv.AddConstraint({"kind": kind, "metadata": {"name": f"c{i}"}, "spec": {}})
    resp = drv.ReviewBatch([], k8s.AUDIT_EP)
    assert resp.n_objects == 0 and resp.results == []
    resp = drv.ReviewBatch([D.Review(object=ok)] * 3, k8s.AUDIT_EP)
    assert resp.viol_bits.shape == (3, 2)
    want = {k for i, k in enumerate(resp.constraints) if k.startswith("NeverValidate/")}
    assert {c for (o, c) in resp.pairs() if o == 1} == want and len(want) == 26
    # removal keeps indices consistent
    drv.RemoveConstraint({"kind": "NeverValidate", "metadata": {"name": "c1"}})
    resp = drv.ReviewBatch([D.Review(object=ok)], k8s.AUDIT_EP)
    assert len(resp.results) == 25


def case_unsupported_is_an_error_not_a_fallback(lib):
    """Constructs outside the lowered subset fail AddTemplate/AddConstraint loudly (no CPU fallback)."""
    import pytest
    t = golden("templates.json")
    drv = D.Driver(lib_path=lib)
    with pytest.raises(D.GkError, match="unsafe"):
        drv.add_template("CompileError", t["fixtures_TemplateCompileError"]["rego"])
    # (referential templates -- data.inventory -- are accepted since round 2: case_referential)
    drv.add_template("K8sUniqueLabel", 'package u\nviolation[{"msg": msg}] {\n  other := data.inventory.cluster[_][_][_]\n'
                     '  other.metadata.labels.x == input.review.object.metadata.labels.x\n  msg := "dup"\n}\n')
    # a pattern taken from the OBJECT (not from the parameters) cannot become a feature column: rejected at AddConstraint
    drv.add_template("PatternFromObject", 'package p\nviolation[{"msg": "m"}] {\n  re_match(input.review.object.metadata.annotations.pat, '
                     'input.parameters.v)\n}\n')
    with pytest.raises(D.GkError, match="rego_unsupported: regular expression"):
        drv.AddConstraint({"kind": "PatternFromObject", "metadata": {"name": "x"}, "spec": {"parameters": {"v": "abc"}}})
    assert drv.Name() == "Rego"
    assert "kernel" in drv.GetDescriptionForStat("kernelTimeNS")


# ------------------------------------------------------------------------------------------ audit aggregation
def case_audit(lib, n=1500, limit=20, excluded=("kube-*", "ns-000*")):
    """The audit sweep's aggregation: excluder stage, totalViolations per constraint and per
    enforcement action, K-smallest status violations in the order updateConstraintStatus emits them, truncation."""
    from oracle import audit as OA
    tm, cons = W.config2()
    # one constraint with a message far beyond 256 bytes exercises truncateString on real results
    long_labels = ["label-%02d-%s" % (i, "x" * 12) for i in range(24)]
    cons = cons + [W._constraint("K8sRequiredLabels", "many-required-labels", params={"labels": long_labels}, action="warn")]
    nss = W.synth_namespaces()
    orc, drv, _ = make_pair(tm, cons, nss, lib_path=lib)
    blob = W.synth_objects(0, n)
    objs = [json.loads(blob.get(i)) for i in range(n)]
    want = OA.audit(orc, objs, namespaces={x["metadata"]["name"]: x for x in nss}, excluded_namespaces=excluded, limit=limit)
    drv.SetExcludedNamespaces("audit", list(excluded))
    run = D.AuditRun(drv, violations_limit=limit)
    half = n // 2
    # two pages, as the audit loop feeds LIST pages (pkg/audit/manager.go:502-561)
    keep = []
    for lo, hi in ((0, half), (half, n)):
        rb = drv.upload([D.Review(object=o, source="Original") for o in objs[lo:hi]], process="audit")
        keep.append(rb)
        run.add_batch(rb, k8s.AUDIT_EP)
    got = run.report()
    want_totals = {"%s/%s" % k: v for k, v in want["totals"].items()}
    assert got["totalViolations"] == want_totals
    assert got["totalViolationsPerEnforcementAction"] == want["by_action"]
    assert sum(want_totals.values()) == got["results"] > 0
    assert any(len(v["message"]) == 256 and v["message"].endswith("...") for v in got["violations"]["K8sRequiredLabels/many-required-labels"])
    for key, lst in want["violations"].items():
        g = got["violations"]["%s/%s" % key]
        w = [{k: v for k, v in sv.items() if not (k in ("namespace", "enforcementActions") and not v)} for sv in lst]
        assert g == w, (key, g[:2], w[:2])
        assert len(g) <= limit
    # without the process flag nothing is excluded
    rb = drv.upload([D.Review(object=o, source="Original") for o in objs[:200]])
    run2 = D.AuditRun(drv, violations_limit=3)
    run2.add_batch(rb, k8s.AUDIT_EP)
    want2 = OA.audit(orc, objs[:200], namespaces={x["metadata"]["name"]: x for x in nss}, limit=3)
    assert run2.report()["totalViolations"] == {"%s/%s" % k: v for k, v in want2["totals"].items()}
    return got


def case_audit_expansion(lib, n_pods=300, limit=5):
    """Expansion inside the audit loop (pkg/audit/manager.go:733-765) through the audit API -- resident batches + gk_audit_add_batch, host-
    flattened and raw-JSON pages: every object is expanded, the resultants' results are the parent's ("[Implied by ...]", the action
    override, the parent's identity in the status violation), an object whose expansion fails contributes nothing and is reported,
    the excluder stage applies to parents and resultants alike.  Report == oracle/audit.py with an expansion system."""
    from oracle import audit as OA
    from oracle import expansion as X
    t = golden("templates.json")
    tm = [(t[k]["kind"], t[k]["rego"]) for k in ("allowedrepos", "psp_privileged", "requiredlabels_basic")]
    cons = [W._constraint(t["allowedrepos"]["kind"], "repos", match=dict(W.POD), params={"repos": ["gcr.io/"]}, action="warn"),
            W._constraint(t["psp_privileged"]["kind"], "priv", match={"kinds": [{"apiGroups": [""], "kinds": ["Pod"]}], "source": "Generated"}),
            W._constraint(t["psp_privileged"]["kind"], "priv-scoped", match=dict(W.POD), action="scoped",
                          scoped=[{"action": "warn", "enforcementPoints": [{"name": k8s.AUDIT_EP}]}]),
            W._constraint(t["requiredlabels_basic"]["kind"], "team", match={"kinds": [{"apiGroups": ["apps"], "kinds": ["Deployment"]}]}, params={"labels": ["team"]})]
    nss = W.synth_namespaces()
    orc, drv, _ = make_pair(tm, cons, nss, lib_path=lib)
    tmpl = lambda name, kinds, gen, src="spec.template", action=None, groups=("apps",): {
        "apiVersion": "expansion.gatekeeper.sh/v1alpha1", "kind": "ExpansionTemplate", "metadata": {"name": name},
        "spec": dict({"applyTo": [{"groups": list(groups), "versions": ["v1"], "kinds": kinds}], "templateSource": src,
                      "generatedGVK": {"group": gen[0], "version": gen[1], "kind": gen[2]}}, **({"enforcementAction": action} if action else {}))}
    xs = X.System()
    for d in (tmpl("expand-deployments", ["Deployment", "ReplicaSet"], ("", "v1", "Pod")),
              tmpl("expand-cronjobs", ["CronJob"], ("batch", "v1", "Job"), src="spec.jobTemplate", groups=("batch",)),
              tmpl("expand-jobs", ["Job"], ("", "v1", "Pod"), action="dryrun", groups=("batch",))):
        xs.upsert(d)
        drv.AddExpansionTemplate(d)
    nsnames = [x["metadata"]["name"] for x in nss]
    bad = {"metadata": {"labels": {"app": "x"}}, "spec": {"containers": [{"name": "c", "image": "evil.example.com/a:latest", "securityContext": {"privileged": True}}]}}
    good = {"metadata": {"labels": {"app": "y"}}, "spec": {"containers": [{"name": "c", "image": "gcr.io/a:1"}]}}
    blob = W.synth_objects(4242, n_pods)
    objs = [json.loads(blob.get(i)) for i in range(n_pods)]
    rnd = random.Random(9)
    for i in range(60):
        ns = rnd.choice(nsnames + ["kube-system", "unknown-ns"])
        spec = rnd.choice([bad, bad, good])
        kind = rnd.choice(["Deployment", "Deployment", "ReplicaSet", "CronJob"])
        if kind == "CronJob":
            o = {"apiVersion": "batch/v1", "kind": "CronJob", "metadata": {"name": "cj-%d" % i, "namespace": ns}, "spec": {"jobTemplate": {"spec": {"template": spec}}}}
        else:
            o = {"apiVersion": "apps/v1", "kind": kind, "metadata": {"name": "%s-%d" % (kind.lower(), i), "namespace": ns, **({"labels": {"team": "a"}} if i % 3 else {})},
                 "spec": {"replicas": 1, "template": spec}}
        objs.append(o)
    # generators whose expansion fails: no spec.template at all, a template that is not a map, a namespace that is not a string
    objs.append({"apiVersion": "apps/v1", "kind": "Deployment", "metadata": {"name": "no-template", "namespace": nsnames[0]}, "spec": {"replicas": 1}})
    objs.append({"apiVersion": "apps/v1", "kind": "Deployment", "metadata": {"name": "bad-template", "namespace": nsnames[0]}, "spec": {"template": "nope"}})
    rnd.shuffle(objs)
    nsmap = {x["metadata"]["name"]: x for x in nss}
    excluded = ("kube-*",)
    want = OA.audit(orc, objs, namespaces=nsmap, excluded_namespaces=excluded, limit=limit, expansion=xs)
    assert len(want["expand_errors"]) == 2, want["expand_errors"]
    assert any("[Implied by expand-jobs]" in r["msg"] and r["enforcementAction"] == "dryrun" for r in want["results"])
    assert any("[Implied by expand-deployments]" in r["msg"] and r["constraint"][1] == "priv-scoped" and r["scopedEnforcementActions"] for r in want["results"])
    drv.SetExcludedNamespaces("audit", list(excluded))

    def check(run):
        got = run.report()
        assert got["totalViolations"] == {"%s/%s" % k: v for k, v in want["totals"].items()}, (got["totalViolations"], want["totals"])
        assert got["totalViolationsPerEnforcementAction"] == want["by_action"]
        for key, lst in want["violations"].items():
            g = got["violations"]["%s/%s" % key]
            w = [{k: v for k, v in sv.items() if not (k in ("namespace", "enforcementActions") and not v)} for sv in lst]
            assert g == w, (key, g[:2], w[:2])
        assert got["objectErrors"]["count"] == len(want["expand_errors"])
        assert sorted(x["error"] for x in got["objectErrors"]["first"]) == sorted(want["expand_errors"].values())
        return got

    keep = []
    run = D.AuditRun(drv, violations_limit=limit)
    half = len(objs) // 2
    for lo, hi in ((0, half), (half, len(objs))):
        rb = drv.upload([D.Review(object=o, source="Original") for o in objs[lo:hi]], process="audit")
        keep.append(rb)
        run.add_batch(rb, k8s.AUDIT_EP)
    got = check(run)
    run = D.AuditRun(drv, violations_limit=limit)
    pb = W.PyBlob([json.dumps(o).encode() for o in objs])
    rb = drv.upload_blob(pb, process="audit")
    keep.append((pb, rb))
    run.add_batch(rb, k8s.AUDIT_EP)
    check(run)
    # the bitmap-only entry points do not expand: with templates registered they refuse instead of returning incomplete rows
    for call in (lambda: rb.eval(k8s.AUDIT_EP), lambda: drv.ReviewBlob(pb, k8s.AUDIT_EP, with_results=False)):
        try:
            call()
            raise AssertionError("a bitmap-only entry point accepted a page while ExpansionTemplates are registered")
        except D.GkError as e:
            assert "ExpansionTemplates are registered" in str(e)
    return got


def case_validation_messages(lib):
    """getValidationMessages (a-14): "[<constraint name>] <msg>" lists per admission request, scoped actions resolved for
    the webhook enforcement point, excluder stage of the webhook process."""
    psp = golden("psp_suite.json")
    cons = [json.loads(json.dumps(c)) for c in psp["constraints"]]
    cons[0].setdefault("spec", {})["enforcementAction"] = "warn"
    cons[1]["spec"]["enforcementAction"] = "dryrun"
    cons[2]["spec"]["enforcementAction"] = "scoped"
    cons[2]["spec"]["scopedEnforcementActions"] = [
        {"action": "warn", "enforcementPoints": [{"name": k8s.WEBHOOK_EP}]},
        {"action": "deny", "enforcementPoints": [{"name": "*"}]},
        {"action": "dryrun", "enforcementPoints": [{"name": k8s.AUDIT_EP}]}]
    cons[3]["spec"]["enforcementAction"] = "scoped"
    cons[3]["spec"]["scopedEnforcementActions"] = [{"action": "deny", "enforcementPoints": [{"name": k8s.AUDIT_EP}]}]
    orc, drv, _ = make_pair([(t["kind"], t["rego"]) for t in psp["templates"]], cons, lib_path=lib)
    pods = [json.loads(json.dumps(p)) for p in psp["pods"]]
    pods[-1]["metadata"]["namespace"] = "kube-system"
    revs = [D.Review(object=p, old_object=p, operation="UPDATE", namespace_name=p["metadata"].get("namespace")) for p in pods]
    drv.SetExcludedNamespaces("webhook", ["kube-*"])
    got = drv.ValidationMessages(revs, process="webhook")
    any_deny = any_warn = False
    for i, r in enumerate(revs):
        if k8s.is_namespace_excluded(["kube-*"], pods[i]):
            assert got[i] == ([], [])
            continue
        rv = k8s.Review(obj=r.object, old=r.old_object, operation="UPDATE", namespace=r.namespace_name)
        res = sorted(orc.review(rv, k8s.WEBHOOK_EP), key=lambda x: (x["constraint"], x["msg"]))
        deny, warn = k8s.validation_messages(res)
        assert (sorted(got[i][0]), sorted(got[i][1])) == (sorted(deny), sorted(warn))
        any_deny, any_warn = any_deny or bool(deny), any_warn or bool(warn)
    assert any_deny and any_warn
    return got


# ------------------------------------------------------------------------------------------ mutation fuzz
_WEIRD = [None, True, False, 0, -1, 1, 65535, 65536, 2 ** 31, 2 ** 53 + 1, 0.5, 1e30, "", "0", "1", "true", "latest", "500m", "0.5", "1Gi", "1e3", "é", "x" * 31,
          "x" * 32, "x" * 300, "a:b:c", ":", "/", "//foo/", "registry.k8s.io/", "gcr.io/proj-01/img@sha256:" + "0" * 64, [], [None], ["a"], {}, {"a": 1},
          [[1]], {"privileged": "true"}, 9223372036854775807, -9223372036854775808, 9223372036854775808]


def _mutate(rnd, doc, n_mut):
    """Random structural damage: delete / retype / replace / duplicate members anywhere in the object."""
    def nodes(x, path, out):
        if isinstance(x, dict):
            for k, v in x.items():
                out.append((x, k))
                nodes(v, path + [k], out)
        elif isinstance(x, list):
            for i, v in enumerate(x):
                out.append((x, i))
                nodes(v, path + [i], out)
    for _ in range(n_mut):
        sites = []
        nodes(doc, [], sites)
        # never damage what makes the document reviewable at all (kind / apiVersion): those are request-level errors,
        # covered by case_review_errors
        sites = [(c, k) for c, k in sites if not (c is doc and k in ("kind", "apiVersion"))]
        if not sites:
            break
        cont, key = rnd.choice(sites)
        op = rnd.random()
        if op < 0.25:
            if isinstance(cont, dict):
                del cont[key]
            else:
                cont.pop(key)
        elif op < 0.75:
            cont[key] = json.loads(json.dumps(rnd.choice(_WEIRD)))
        elif op < 0.85 and isinstance(cont, list):
            cont.append(json.loads(json.dumps(cont[key])))
        elif isinstance(cont, dict):
            cont[rnd.choice(["name", "image", "hostPort", "privileged", "cpu", "memory", "team", "path", "readOnly", "extra"])] = json.loads(
                json.dumps(rnd.choice(_WEIRD)))
    return doc


def case_fuzz(lib, n=800, seed=1234, start=5000):
    """config-2 constraints against synthetic Pods with random structural damage: wrong types, nulls, missing members,
    out-of-range numbers, strings on the 31/32/255-byte boundaries of the column encodings.  Objects whose numbers the
    device cannot compare exactly are per-object review errors in the engine and are excluded from the comparison."""
    import random
    rnd = random.Random(seed)
    tm, cons = W.config2()
    nss = W.synth_namespaces()
    orc, drv, _ = make_pair(tm, cons, nss, lib_path=lib)
    blob = W.synth_objects(start, n)
    objs = []
    for i in range(n):
        o = json.loads(blob.get(i))
        objs.append(_mutate(rnd, o, rnd.choice([0, 1, 1, 2, 3, 5])))
    revs = []
    for i, o in enumerate(objs):
        r = rnd.random()
        if r < 0.15 and i:       # UPDATE carrying another (damaged) Pod as oldObject: either may match (matcher.go:44-71)
            revs.append(D.Review(object=o, old_object=objs[i - 1], operation="UPDATE", source="Original"))
        elif r < 0.20:           # DELETE: the object under review IS the old object (target.go:262-280)
            revs.append(D.Review(object=None, old_object=o, operation="DELETE", source="Original"))
        else:
            revs.append(D.Review(object=o, source="Original"))
    resp = drv.ReviewBatch(revs, k8s.AUDIT_EP)
    errs = resp.object_errors or [None] * n
    bad = {i for i, e in enumerate(errs) if e}
    assert len(bad) < n // 4, "too many objects refused: %r" % [errs[i] for i in sorted(bad)[:5]]
    for i in bad:
        assert "int64" in errs[i] or "invalid request object" in errs[i], errs[i]
    want = set()
    for x in oracle_results_safe(orc, revs, k8s.AUDIT_EP, skip=bad):
        want.add(x)
    got = {x for x in engine_results(resp) if x[0] not in bad}
    assert_same(want, got)
    return len(want), len(bad)


def oracle_results_safe(orc, reviews, ep, skip=()):
    """oracle_results for the objects not in `skip` (keeps the original object indices)."""
    out = set()
    for i, r in enumerate(reviews):
        if i in skip:
            continue
        for x in oracle_results(orc, [r], ep):
            out.add((i,) + x[1:])
    return out


# ------------------------------------------------------------------------------------------ match fuzz
def case_match_fuzz(lib, n_constraints=60, n_objects=400, seed=5, extra_names=()):
    """Random `spec.match` blocks x random review shapes through the in-kernel pre-filter, against the oracle's
    restatement of match.Matches / Matcher.Match (which the reference's own vectors pin).  The template always
    violates, so the result set is exactly the match relation (plus autoreject results for matcher errors)."""
    import random
    rnd = random.Random(seed)
    t = golden("templates.json")["fixtures_TemplateNeverValidate"]
    names = ["a", "ab", "abc", "kube-system", "kube-public", "prod-1", "prod-22", "dev", "x-system"] + list(extra_names) + [""]
    wild = lambda: rnd.choice(names[:-1] + ["*", "a*", "*a", "*b*", "kube-*", "*-system", "prod-*", "*-1", "**"])
    keys, vals = ["app", "team", "tier", "env", "example.com/role", "bad key!", "", "a/b/c", "-x"], ["a", "b", "c", "", "not ok", "x" * 64]
    def selector():
        s = {}
        if rnd.random() < 0.6:
            s["matchLabels"] = {rnd.choice(keys): rnd.choice(vals) for _ in range(rnd.randint(0, 2))}
        if rnd.random() < 0.6:
            s["matchExpressions"] = []
            for _ in range(rnd.randint(0, 2)):
                op = rnd.choice(["In", "NotIn", "Exists", "DoesNotExist"])
                e = {"key": rnd.choice(keys), "operator": op}
                if op in ("In", "NotIn"):
                    e["values"] = [rnd.choice(vals[:3]) for _ in range(rnd.randint(1, 3))]
                elif rnd.random() < 0.1:
                    e["values"] = ["a"]                      # invalid: Exists with values
                s["matchExpressions"].append(e)
        return s
    cons = []
    for i in range(n_constraints):
        m = {}
        if rnd.random() < 0.5:
            m["kinds"] = [{"apiGroups": rnd.choice([[""], ["*"], ["apps"], ["", "apps"], []]),
                           "kinds": rnd.choice([["Pod"], ["*"], ["Deployment", "Pod"], ["Namespace"], []])} for _ in range(rnd.randint(1, 2))]
        if rnd.random() < 0.3:
            m["scope"] = rnd.choice(["Cluster", "Namespaced", "*", "Bogus"])
        if rnd.random() < 0.4:
            m["namespaces"] = [wild() for _ in range(rnd.randint(1, 3))]
        if rnd.random() < 0.4:
            m["excludedNamespaces"] = [wild() for _ in range(rnd.randint(1, 3))]
        if rnd.random() < 0.4:
            m["labelSelector"] = selector()
        if rnd.random() < 0.4:
            m["namespaceSelector"] = selector()
        if rnd.random() < 0.3:
            m["name"] = wild()
        if rnd.random() < 0.3:
            m["source"] = rnd.choice(["All", "Original", "Generated", "", "Bogus"])
        cons.append({"kind": t["kind"], "metadata": {"name": "m%03d" % i}, "spec": ({"match": m} if (m or rnd.random() < 0.5) else {})})
    cached = [{"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": n, "labels": {rnd.choice(keys): rnd.choice(vals) for _ in range(rnd.randint(0, 3))}}}
              for n in names[:6]]
    orc, drv, _ = make_pair([(t["kind"], t["rego"])], cons, cached, lib_path=lib)
    def obj():
        kind, group = rnd.choice([("Pod", ""), ("Pod", ""), ("Deployment", "apps"), ("Namespace", ""), ("ClusterRole", "rbac.authorization.k8s.io")])
        meta = {}
        if rnd.random() < 0.85:
            meta["name"] = rnd.choice(names)
        if rnd.random() < 0.2:
            meta["generateName"] = rnd.choice(names)
        if kind in ("Pod", "Deployment") and rnd.random() < 0.9:
            meta["namespace"] = rnd.choice(names)
        if rnd.random() < 0.7:
            meta["labels"] = {rnd.choice(keys): rnd.choice(vals) for _ in range(rnd.randint(0, 3))}
        return {"apiVersion": (group + "/v1") if group else "v1", "kind": kind, "metadata": meta}
    revs = []
    for _ in range(n_objects):
        o = obj()
        kw = {"object": o, "source": rnd.choice(["Original", "Generated", "", "All"])}
        r = rnd.random()
        if r < 0.15:
            kw.update(old_object=obj(), operation="UPDATE")
        elif r < 0.25:
            kw = {"object": None, "old_object": o, "operation": "DELETE", "source": kw["source"]}
        if rnd.random() < 0.25:
            kw["namespace"] = {"apiVersion": "v1", "kind": "Namespace",
                               "metadata": {"name": o["metadata"].get("namespace", "zz"), "labels": {rnd.choice(keys): rnd.choice(vals)}}}
        revs.append(D.Review(**kw))
    for ep in (k8s.AUDIT_EP, k8s.WEBHOOK_EP):
        resp = drv.ReviewBatch(revs, ep)
        assert_same(oracle_results(orc, revs, ep), engine_results(resp))
    return len(resp.results)


def case_fuzz_other_templates(lib, n=400, seed=77):
    """The in-tree templates outside config 2 (regex labels, object.get defaults, namespaceObject, user info, custom
    fields, the comprehension form of allowed repos, `violation contains ... if`), with varied parameters, against
    mutated objects of several kinds."""
    import random
    rnd = random.Random(seed)
    t = golden("templates.json")
    pick = {"requiredlabels_agilebank": [{"labels": [{"key": "owner", "allowedRegex": "^[a-z]+[.]agilebank[.]demo$"}, {"key": "team"}]},
                                         {"message": "custom message", "labels": [{"key": "team", "allowedRegex": "^(a|b|team-[0-9]+)$"}]},
                                         {"labels": [{"key": "app", "allowedRegex": ""}]}],
            "requiredlabels_regov1": [{"labels": ["team", "owner"]}, {"labels": []}],
            "fooischeck": [{"foo": "bar"}, {"foo": ""}, {}, {"foo": 7}],
            "namespacelabelcheck": [{"requiredLabel": "team"}, {"requiredLabel": "bar"}],
            "allowedrepos": [{"repos": ["gcr.io/", "quay.io/"]}, {"repos": []}, {"repos": ["openpolicyagent/opa:", "docker.io/library/nginx"]}],
            "fixtures_TemplateValidateUserInfo": [None],
            "fixtures_TemplateRestrictCustomField": [{"expectedCustomField": "x"}, {"expectedCustomField": 7}, {"expectedCustomField": {"a": [1, 2]}},
                                                     {"expectedCustomField": None}]}
    tm, cons = [], []
    for name, plist in pick.items():
        tm.append((t[name]["kind"], t[name]["rego"]))
        for i, params in enumerate(plist):
            cons.append(W._constraint(t[name]["kind"], "%s-%d" % (name.replace("_", "-").lower(), i), params=params,
                                      action=rnd.choice([None, "warn", "dryrun"])))
    nss = W.synth_namespaces()
    orc, drv, skipped = make_pair(tm, cons, nss, lib_path=lib, skip_unsupported=True)
    assert not skipped, skipped
    blob = W.synth_objects(31000 + seed, n)
    revs = []
    for i in range(n):
        o = json.loads(blob.get(i))
        if rnd.random() < 0.3:
            o.setdefault("metadata", {}).setdefault("labels", {})[rnd.choice(["owner", "team", "app"])] = rnd.choice(
                ["alice.agilebank.demo", "bob", "team-7", "a", "", "Team-1", "x.agilebank.demo.evil"])
        if rnd.random() < 0.3:
            o["foo"] = rnd.choice(["bar", "", "baz", 7, None, ["bar"]])
        if rnd.random() < 0.3:
            o.setdefault("spec", {})["customField"] = rnd.choice(["x", "y", 7, 7.0, {"a": [1, 2]}, {"a": [2, 1]}, None, [1]])
        o = _mutate(rnd, o, rnd.choice([0, 0, 1, 2]))
        kw = {"object": o, "source": "Original"}
        if rnd.random() < 0.5:
            kw["user_info"] = {"username": rnd.choice(["system:serviceaccount:kube-system:x", "alice", "", "system:", "System:admin"])}
        if rnd.random() < 0.3:
            kw["namespace"] = {"apiVersion": "v1", "kind": "Namespace",
                               "metadata": dict({"name": rnd.choice(["explicit-ns", "ns-0001"])},
                                                **({"labels": {rnd.choice(["team", "bar", "other"]): "v"}} if rnd.random() < 0.7 else {}))}
        revs.append(D.Review(**kw))
    n_results = 0
    for ep in (k8s.AUDIT_EP, k8s.WEBHOOK_EP):
        resp = drv.ReviewBatch(revs, ep)
        errs = resp.object_errors or [None] * n
        bad = {i for i, e in enumerate(errs) if e}
        assert len(bad) < n // 4
        want = oracle_results_safe(orc, revs, ep, skip=bad)
        got = {x for x in engine_results(resp) if x[0] not in bad}
        assert_same(want, got)
        n_results += len(want)
    return n_results


def case_template_libs(lib, n=300, seed=99):
    """A template with `spec.targets[].libs` (test/bats/tests/templates/k8scontainterlimits_template.yaml: the entry point
    imports data.lib.helpers): the bats expectation (test/bats/test.bats:268-279 -- opa_no_limits.yaml denied, opa.yaml
    admitted), then fuzzed limits / parameters against the oracle."""
    import random
    g = golden("libs_template.json")
    tmpl = k8s.template_from_yaml_obj(g["template"])
    assert len(tmpl) == 3 and len(tmpl[2]) == 1
    orc, drv, skipped = make_pair([tmpl], [g["constraint"]], lib_path=lib)
    assert not skipped
    revs = [D.Review(object=g["denied"]), D.Review(object=g["admitted"])]
    resp = drv.ReviewBatch(revs, k8s.WEBHOOK_EP)
    assert_same(oracle_results(orc, revs, k8s.WEBHOOK_EP), engine_results(resp))
    by_obj = {}
    for r in resp.results:
        by_obj.setdefault(r.object, []).append(r.msg)
    assert by_obj.get(0) == ["container <opa> has no resource limits"] and 1 not in by_obj

    rnd = random.Random(seed)
    kind = tmpl[0]
    cons = [W._constraint(kind, "limits-%d" % i, params=p, action=rnd.choice([None, "warn"]))
            for i, p in enumerate([{"cpu": "200m", "memory": "1Gi"}, {"cpu": "2", "memory": "512Mi"}, {"cpu": 1, "memory": 1073741824},
                                   {"cpu": "500m", "memory": "2G"}, {"cpu": "bogus", "memory": "1Ti"}, {"memory": "100M"}, {}])]
    orc, drv, skipped = make_pair([tmpl], cons, lib_path=lib)
    assert not skipped
    cpus = ["100m", "200m", "201m", "1", "2", "3", 1, 2, 0.5, "0.5", "abc", "", "1000m", "m", None]
    mems = ["1Gi", "2Gi", "512Mi", "513Mi", "100M", "101M", "1G", "3G", 1073741824, 2000000000, "1Ki", "1000", "1Ei", "1.5Gi", "Gi", "", "xyz",
            "128974848", "129e6", "1E", "1P", "1T", "100k", "100m", None]
    revs = []
    for i in range(n):
        cs = []
        for j in range(rnd.choice([0, 1, 1, 2, 3])):
            c = {"name": "c%d" % j, "image": "img"}
            mode = rnd.random()
            if mode < 0.1:
                pass
            elif mode < 0.2:
                c["resources"] = rnd.choice([{}, {"requests": {"cpu": "1"}}, None])
            else:
                lim = {}
                cpu, mem = rnd.choice(cpus), rnd.choice(mems)
                if rnd.random() < 0.9:
                    lim["cpu"] = cpu
                if rnd.random() < 0.9:
                    lim["memory"] = mem
                c["resources"] = {"limits": lim}
            cs.append(c)
        revs.append(D.Review(object={"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "p%d" % i, "namespace": "default"},
                                     "spec": {"containers": cs}}))
    resp = drv.ReviewBatch(revs, k8s.AUDIT_EP)
    # "1E" / "1Ei" canonify to 1e21 / 1.15e21 millibytes: beyond the exact int64 columns, reported per object (never guessed)
    bad = {i for i, e in enumerate(resp.object_errors or []) if e}
    assert all("int64" in resp.object_errors[i] for i in bad) and len(bad) < n // 5
    want = oracle_results_safe(orc, revs, k8s.AUDIT_EP, skip=bad)
    assert len(want) > n
    assert_same(want, {x for x in engine_results(resp) if x[0] not in bad})

    # a lib outside `package lib...` and an import of a lib the template does not have are compile errors in both
    import pytest
    from oracle import rego as orego
    main = 'package x\nimport data.lib.nothere\nviolation[{"msg": "m"}] { nothere.f(1) }\n'
    for libs, src in (([("package helpers\nf(x) = x { true }\n")], tmpl[1]), ([], main)):
        with pytest.raises(orego.RegoError, match="rego_compile_error"):
            k8s.Client().add_template("X", src, libs)
        with pytest.raises(D.GkError, match="rego_compile_error"):
            D.Driver(lib_path=lib).add_template("X", src, libs or ["package lib.other\ng(x) = x { true }\n"])
    # libs that import each other, full data.lib paths, and a non-function lib rule
    a = 'package lib.a\nimport data.lib.b\nbig(x) { b.limit < x }\nnames[n] { n := input.review.object.spec.containers[_].name }\n'
    b = 'package lib.b\nlimit = 2 { true }\n'
    main = ('package y\nimport data.lib.a as util\n'
            'violation[{"msg": msg}] { util.big(count(input.review.object.spec.containers)); msg := sprintf("too many: %v", [data.lib.a.names]) }\n'
            'violation[{"msg": msg}] { util.names["c0"]; data.lib.b.limit == 2; msg := "has c0" }\n')
    orc, drv, skipped = make_pair([("Y", main, (a, b))], [W._constraint("Y", "y", params=None)], lib_path=lib)
    assert not skipped
    resp = drv.ReviewBatch(revs, k8s.AUDIT_EP)
    assert not any(resp.object_errors or [])
    want = oracle_results(orc, revs, k8s.AUDIT_EP)
    assert len(want) > 10
    assert_same(want, engine_results(resp))
    return len(want)


def case_more_builtins(lib):
    """Collection / object / rounding builtins (sort, object.keys|union|remove|filter, numbers.range, array.slice|reverse,
    strings.reverse, round|floor|ceil, format_int, union, intersection, product, type_name, base64.*) on object values (host
    feature columns), on parameters (folded at AddConstraint) and compared across the two (device atoms)."""
    src = '''package bx
violation[{"msg": msg}] {
  o := input.review.object
  ks := sort(object.keys(o.metadata.labels))
  msg := sprintf("%v|%v|%v|%v|%v|%v|%v|%v|%v|%v|%v|%v|%v|%v", [ks, object.union(o.a, o.b), object.remove(o.a, ["x"]), object.filter(o.a, {"x"}),
     numbers.range(o.lo, o.hi), array.slice(o.arr, 1, 3), array.reverse(o.arr), strings.reverse(o.s), round(o.f), floor(o.f), ceil(o.f),
     format_int(o.f, 16), type_name(o.arr), base64.decode(base64.encode(o.s))])
}
violation[{"msg": msg}] {
  o := input.review.object
  u := union({{x | x := o.arr[_]}, {1, 99}})
  i := intersection({{x | x := o.arr[_]}, {1, 3, 99}})
  msg := sprintf("sets %v %v %v", [u, i, product(o.arr)])
}
violation[{"msg": msg}] {
  allowed := sort(input.parameters.names)
  first := allowed[0]
  input.review.object.metadata.name == first
  msg := sprintf("first of %v", [allowed])
}
violation[{"msg": msg}] {
  floor(input.review.object.f) > count(numbers.range(1, input.parameters.n))
  msg := sprintf("f above %v", [format_int(input.parameters.n, 2)])
}
'''
    objs = [{"apiVersion": "v1", "kind": "X", "metadata": {"name": "o%d" % i, "labels": {"b": "1", "a": "2", "c": "3"}},
             "a": {"x": 1, "y": {"p": 1, "q": 2}}, "b": {"y": {"q": 3, "r": 4}, "z": 5}, "lo": lo, "hi": hi, "arr": [3, 1, 2, 7], "s": s, "f": f}
            for i, (lo, hi, s, f) in enumerate([(1, 4, "h\u00e9llo", 2.5), (4, 1, "", -2.5), (0, 0, "abc", 3.49), (2, 2, "x", -0.5),
                                                (1, 3, "\u65e5\u672c\u8a9e", 1000.5), (1, 2, "q", 7)])]
    cons = [W._constraint("BX", "bx-%d" % i, params=p) for i, p in enumerate([{"names": ["o3", "o1", "o2"], "n": 3}, {"names": ["zz", "o0"], "n": 2.0},
                                                                                {"names": [], "n": 0}])]
    orc, drv, skipped = make_pair([("BX", src)], cons, lib_path=lib)
    assert not skipped
    revs = [D.Review(object=o) for o in objs]
    resp = drv.ReviewBatch(revs, k8s.AUDIT_EP)
    assert not any(resp.object_errors or [])
    want = oracle_results(orc, revs, k8s.AUDIT_EP)
    assert len(want) >= 3 * 2 * len(objs)
    msgs = {w[2] for w in want}
    assert '["a", "b", "c"]|{"x": 1, "y": {"p": 1, "q": 3, "r": 4}, "z": 5}|{"y": {"p": 1, "q": 2}}|{"x": 1}|[1, 2, 3, 4]|[1, 2]|[7, 2, 1, 3]|oll\u00e9h|3|2|3|2|array|h\u00e9llo' in msgs
    assert 'first of ["o1", "o2", "o3"]' in msgs and "f above 11" in msgs
    assert_same(want, engine_results(resp))
    return len(want)


def case_every(lib, n=800):
    """`every` (OPA v1 keyword): the engine rewrites it into a counted comprehension over a generated helper function, the
    oracle evaluates it natively -- over object collections, parameter lists, with key and value, with an empty and with an
    undefined domain."""
    src = '''package e
violation[{"msg": msg}] {
  every c in input.review.object.spec.containers { startswith(c.image, input.parameters.prefix) }
  msg := "every image has the prefix"
}
violation[{"msg": msg}] {
  c := input.review.object.spec.containers[_]
  every p in input.parameters.prefixes { not startswith(c.image, p) }
  msg := sprintf("container <%v> matches no prefix", [c.name])
}
violation[{"msg": msg}] {
  c := input.review.object.spec.containers[_]
  every i, m in c.volumeMounts { m.readOnly; i < 2 }
  msg := sprintf("container <%v> mounts (at most two) read-only only", [c.name])
}
violation[{"msg": msg}] {
  every k, v in input.review.object.metadata.labels { startswith(k, "label-"); count(v) > input.parameters.n }
  msg := "all labels are label-*"
}
'''
    blob = W.synth_objects(7, n)
    revs = [D.Review(object=json.loads(blob.get(i))) for i in range(n)]
    revs.append(D.Review(object={"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "nolabels"}, "spec": {}}))
    revs.append(D.Review(object={"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "empty", "labels": {}}, "spec": {"containers": []}}))
    cons = [W._constraint("E", "e1", params={"prefix": "gcr.io/", "prefixes": ["gcr.io/", "quay.io/"], "n": 1}),
            W._constraint("E", "e2", params={"prefix": "", "prefixes": [], "n": 3})]
    orc, drv, skipped = make_pair([("E", src)], cons, lib_path=lib)
    assert not skipped
    resp = drv.ReviewBatch(revs, k8s.AUDIT_EP)
    want = oracle_results(orc, revs, k8s.AUDIT_EP)
    by_obj = {}
    for w in want:
        by_obj.setdefault(w[0], set()).add((w[1], w[2]))
    # undefined domain (no labels, no containers): no result; empty domains: vacuously true
    assert n not in by_obj
    assert by_obj[n + 1] == {("E/e1", "every image has the prefix"), ("E/e2", "every image has the prefix"),
                             ("E/e1", "all labels are label-*"), ("E/e2", "all labels are label-*")}
    assert len(want) > n
    assert_same(want, engine_results(resp))
    return len(want)


def case_validate_constraint(lib):
    """TestValidateConstraint (pkg/target/target_test.go:42-399, 11 cases): which constraints ValidateConstraint refuses -- oracle
    and engine -- plus hand-made selectors whose error TEXT the two restatements must agree on."""
    import pytest
    g = golden("validate_constraint_vectors.json")
    drv = D.Driver(lib_path=lib)
    assert len(g["cases"]) == 11
    for cse in g["cases"]:
        oerr = eerr = None
        try:
            k8s.validate_constraint(cse["constraint"])
        except k8s.ValidateError as e:
            oerr = str(e)
        try:
            drv.ValidateConstraint(cse["constraint"])
        except D.GkError as e:
            eerr = str(e)
        assert (oerr is not None) == cse["error_expected"], (cse["name"], oerr)
        assert (eerr is not None) == cse["error_expected"], (cse["name"], eerr)
        if oerr is not None:
            assert oerr in eerr, (cse["name"], oerr, eerr)

    def con(match):
        return {"apiVersion": "constraints.gatekeeper.sh/v1beta1", "kind": "K", "metadata": {"name": "k"}, "spec": {"match": match}}
    extra = [{"labelSelector": {"matchLabels": {"bad key!": "v", "ok": "bad value!"}}},
             {"labelSelector": {"matchExpressions": [{"key": "a", "operator": "In", "values": []}, {"key": "b", "operator": "Exists", "values": ["x"]}]}},
             {"namespaceSelector": {"matchExpressions": [{"key": "", "operator": "NotIn", "values": ["not ok", "x" * 64]}]}},
             {"labelSelector": {"matchExpressions": [{"key": "a/b/c", "operator": "DoesNotExist"}]}},
             {"labelSelector": {"matchExpressions": {"key": "a"}}}, {"labelSelector": {"matchExpressions": [{"key": 1, "operator": "In"}]}},
             {"labelSelector": {"matchExpressions": [{"key": "a", "operator": "In", "values": [1]}]}}, {"labelSelector": None, "namespaceSelector": {}},
             {"labelSelector": {"matchLabels": {"a": 1}}}, {"kinds": [{"kinds": ["Pod"]}]}]
    n_err = 0
    for m in extra:
        oerr = eerr = None
        try:
            k8s.validate_constraint(con(m))
        except k8s.ValidateError as e:
            oerr = str(e)
        try:
            drv.ValidateConstraint(con(m))
        except D.GkError as e:
            eerr = str(e)
        assert (oerr is None) == (eerr is None), (m, oerr, eerr)
        if oerr is not None:
            n_err += 1
            assert oerr in eerr, (m, oerr, eerr)
    assert n_err == 8
    # ToMatcher (pkg/target/target.go:239-254) runs at AddConstraint: TestToMatcher's two invalid constraints (target_test.go:544-560:
    # spec.match = 3.0, spec.match.kinds = 3.0 => ErrCreatingMatcher) and other members of the wrong JSON type
    rego_src = 'package k\nviolation[{"msg": "m"}] { true }\n'
    n_match_err = 0
    for m in (3.0, {"kinds": 3.0}, {"namespaces": "x"}, {"labelSelector": {"matchLabels": 3}}, {"name": 7}, {"kinds": [{"apiGroups": [1]}]},
              {"scope": []}, None, {"kinds": [{"apiGroups": [""], "kinds": ["Pod"]}], "namespaces": ["a*"], "name": "x*", "scope": "Namespaced"}):
        c = con({}) if m is None else con(m)
        if m is None:
            del c["spec"]["match"]
        o, d = k8s.Client(), D.Driver(lib_path=lib)
        o.add_template("K", rego_src)
        d.add_template("K", rego_src)
        oerr = eerr = None
        try:
            o.add_constraint(c)
        except k8s.MatchError as e:
            oerr = str(e)
        try:
            d.AddConstraint(c)
        except D.GkError as e:
            eerr = str(e)
        assert (oerr is None) == (eerr is None), (m, oerr, eerr)
        if oerr is not None:
            n_match_err += 1
            assert oerr.startswith("unable to create matcher: ") and oerr in eerr, (m, oerr, eerr)
    assert n_match_err == 7
    # Beyond that, AddConstraint is the DRIVER's method and does not validate (the reference's client validates first): the matcher
    # vectors rely on invalid selectors reaching Matches()
    return n_err


# ------------------------------------------------------------------------------------------ random policies
_RF_HELPERS = """
input_containers[c] { c := input.review.object.spec.containers[_] }
input_containers[c] { c := input.review.object.spec.initContainers[_] }
has_probe(c) { c.readinessProbe }
has_probe(c) { c.livenessProbe }
tag_of(image) = t { parts := split(image, ":"); count(parts) > 1; t := parts[count(parts) - 1] }
tag_of(image) = "latest" { not contains(image, ":") }
level(c) = "high" { c.securityContext.privileged } else = "low" { true }
every_mount_under(c, prefix) { every m in c.volumeMounts { startswith(m.mountPath, prefix) } }
"""

_RF_CONDS = {
    "container": [
        'not startswith(c.image, input.parameters.prefix)',
        'startswith(c.image, input.parameters.prefixes[_])',
        'satisfied := [good | repo = input.parameters.prefixes[_]; good = startswith(c.image, repo)]\n  not any(satisfied)',
        'not strings.any_prefix_match(c.image, input.parameters.prefixes)',
        'c.securityContext.privileged',
        'not c.securityContext.privileged',
        'not c.resources.limits.cpu',
        'c.resources.limits.cpu == input.parameters.cpu',
        'c.resources.limits.memory != input.parameters.mem',
        'to_number(c.resources.limits.cpu) > input.parameters.n',
        'endswith(c.image, ":latest")',
        'contains(c.image, input.parameters.sub)',
        'count(c.ports) > input.parameters.n',
        'count(c.volumeMounts) >= input.parameters.n',
        'c.name != input.parameters.name',
        're_match(input.parameters.pattern, c.image)',
        'not re_match("^[a-z0-9./-]+:[a-z0-9.]+$", c.image)',
        'tag_of(c.image) == input.parameters.tag',
        'tag_of(c.image) != "latest"',
        'not has_probe(c)',
        'has_probe(c)',
        'input.parameters.tags[_] == tag_of(c.image)',
        'p := c.ports[_]\n  p.hostPort > input.parameters.n',
        'p := c.ports[_]\n  not p.hostPort',
        'm := c.volumeMounts[_]\n  not m.readOnly\n  startswith(m.mountPath, input.parameters.mount)',
        'object.get(c, "imagePullPolicy", "Always") == input.parameters.policy',
        'lower(c.name) == input.parameters.name',
        # stranger shapes: joins across scopes, object-object compares, arithmetic, more builtins, `in`, `else`
        'm := c.volumeMounts[_]\n  vol := input.review.object.spec.volumes[_]\n  vol.name == m.name\n  vol.hostPath',
        'm := c.volumeMounts[_]\n  vol := input.review.object.spec.volumes[_]\n  vol.name == m.name\n  not m.readOnly\n  not vol.emptyDir',
        'c.resources.limits.cpu == c.resources.requests.cpu',
        'x := to_number(c.resources.limits.cpu) * 1000\n  x > input.parameters.n',
        'replace(c.image, ":", "@") == input.parameters.sub',
        'trim(c.name, "c") == "0"',
        'upper(c.name) == "C0"',
        'concat("/", [input.review.object.metadata.namespace, c.name]) == input.parameters.name',
        'sprintf("%v", [c.name]) == input.parameters.name',
        'count([p | p := c.ports[_]; p.hostPort]) > 0',
        'ports := {p.containerPort | p := c.ports[_]}\n  ports[input.parameters.n]',
        'c.securityContext.privileged == true',
        'is_string(c.resources.limits.cpu)',
        'is_number(c.resources.limits.cpu)',
        'indexof(c.image, "/") > input.parameters.n',
        'substring(c.image, 0, 3) == "gcr"',
        'some p in c.ports\n  p.hostPort == input.parameters.n',
        'tag_of(c.image) in input.parameters.tags',
        'not tag_of(c.image) in input.parameters.tags',
        'level(c) == input.parameters.tag',
        'level(c) != "low"',
        # boolean VALUES (an undefined builtin result is not a false one)
        'x := startswith(c.image, input.parameters.prefix)\n  not x',
        'x := re_match(input.parameters.pattern, c.image)\n  not x',
        'x := c.image == input.parameters.sub\n  not x',
        'x := tag_of(c.image) in input.parameters.tags\n  not x',
        'x := endswith(c.image, ":latest")\n  x == false',
        'x := count(c.ports) > input.parameters.n\n  not x',
        'goods := [g | p := input.parameters.prefixes[_]; g := startswith(c.image, p)]\n  count(goods) == 0',
        'goods := [g | p := input.parameters.prefixes[_]; g := startswith(c.image, p)]\n  all(goods)',
        'goods := [g | p := input.parameters.prefixes[_]; g := endswith(c.image, p)]\n  not all(goods)',
        'goods := [g | p := input.parameters.prefixes[_]; g := contains(c.image, p)]\n  count(goods) > 0',
        'flags := [startswith(c.image, input.parameters.prefix), c.name == input.parameters.name]\n  any(flags)',
        'flags := [startswith(c.image, input.parameters.prefix), c.name != input.parameters.name]\n  all(flags)',
        'count([p | p := c.ports[_]; p.containerPort > input.parameters.n]) > 0',
        'every m in c.volumeMounts { m.readOnly }',
        'not every_mount_under(c, input.parameters.mount)',
        'every p in input.parameters.prefixes { not startswith(c.image, p) }',
        'every i, p in c.ports { p.containerPort > input.parameters.n; i < 2 }',
        'all([m.readOnly | m := c.volumeMounts[_]; startswith(m.mountPath, input.parameters.mount)])',
        'not any([m.readOnly | m := c.volumeMounts[_]; startswith(m.mountPath, input.parameters.mount)])',
    ],
    "label": [
        'k == input.parameters.key',
        'v != input.parameters.val',
        'startswith(k, "label-0")',
        'input.parameters.labels[_] == k',
        'not re_match(input.parameters.pattern, v)',
        'count(v) > input.parameters.n',
        'endswith(v, input.parameters.sub)',
    ],
    "volume": [
        'vol.hostPath',
        'not vol.emptyDir',
        'startswith(vol.hostPath.path, input.parameters.prefix)',
        'fields := {x | vol[x]; x != "name"}\n  count(fields - {y | y := input.parameters.volumes[_]}) > 0',
        'vol.name == input.parameters.name',
        'vol.persistentVolumeClaim.claimName != input.parameters.name',
    ],
    "none": [
        'provided := {l | input.review.object.metadata.labels[l]}\n  required := {l | l := input.parameters.labels[_]}\n  missing := required - provided\n  count(missing) > 0',
        'input.review.object.spec.hostNetwork',
        'not input.review.object.metadata.labels[input.parameters.key]',
        'count(input.review.object.spec.containers) > input.parameters.n',
        'input.review.object.metadata.namespace == input.parameters.ns',
        'object.get(input.review.object.spec, "hostPID", false) == input.parameters.flag',
        'input.review.object.metadata.labels[input.parameters.key] == input.parameters.val',
        'count({c | c := input_containers[_]; c.securityContext.privileged}) >= input.parameters.n',
        'input.review.kind.kind == input.parameters.kind',
        'not input.parameters.flag',
        'input.parameters.n > 1',
        'ns := input.review.object.metadata.namespace\n  not startswith(ns, input.parameters.prefix)',
        'count(input.review.object.spec.containers) != count(input.review.object.spec.volumes)',
        'names := [c.name | c := input.review.object.spec.containers[_]]\n  concat(",", names) == input.parameters.name',
        'input.review.object.spec.containers[0].image == input.parameters.sub',
        'input.review.object.metadata.labels.team',
        'all([startswith(c.image, input.parameters.prefix) | c := input.review.object.spec.containers[_]])',
        'any([c.securityContext.privileged | c := input_containers[_]])',
        'x := input.parameters.n + 1\n  count(input.review.object.spec.containers) < x',
        'input.review.object.metadata.name == input.review.object.spec.containers[_].name',
        'input.review.object.spec.containers[_].name == input.review.object.spec.volumes[_].name',
        'not input.review.object.spec.volumes',
        # structured parameters (lists of objects, as in K8sRequiredLabels' {key, allowedRegex})
        'expected := input.parameters.pairs[_]\n  input.review.object.metadata.labels[expected.key] != expected.val',
        'some pair in input.parameters.pairs\n  not input.review.object.metadata.labels[pair.key]',
        'expected := input.parameters.pairs[_]\n  v := input.review.object.metadata.labels[expected.key]\n  not re_match(expected.pattern, v)',
        'every pair in input.parameters.pairs { input.review.object.metadata.labels[pair.key] }',
        'wanted := {p.key | p := input.parameters.pairs[_]}\n  have := {k | input.review.object.metadata.labels[k]}\n  count(wanted - have) > 0',
        'count({p.key | p := input.parameters.pairs[_]; input.review.object.metadata.labels[p.key] == p.val}) == 0',
        # the review document around the object
        'input.review.operation == "UPDATE"',
        'input.review.operation != "DELETE"',
        'input.review.oldObject.metadata.labels[input.parameters.key] != input.review.object.metadata.labels[input.parameters.key]',
        'not input.review.oldObject.metadata',
        'count(input.review.oldObject.spec.containers) != count(input.review.object.spec.containers)',
        'startswith(input.review.userInfo.username, "system:")',
        'not input.review.userInfo.username',
        'not input.review.namespaceObject.metadata.labels[input.parameters.key]',
        'input.review.namespaceObject.metadata.labels[input.parameters.key] == input.parameters.val',
        'input.review.kind.group == ""',
        'input.review.name == input.review.object.metadata.name',
        'input.review.namespace == input.parameters.ns',
        'every c in input.review.object.spec.containers { startswith(c.image, input.parameters.prefix) }',
        'every c in input_containers { c.resources.limits.cpu; not c.securityContext.privileged }',
        'every k, v in input.review.object.metadata.labels { startswith(k, "label-"); count(v) > input.parameters.n }',
        'every vol in input.review.object.spec.volumes { some t in input.parameters.volumes; vol[t] }',
        'not all([startswith(c.image, input.parameters.prefix) | c := input.review.object.spec.containers[_]])',
        'any([endswith(c.image, input.parameters.sub) | c := input_containers[_]])',
        'count([c | c := input.review.object.spec.containers[_]; startswith(c.image, input.parameters.prefix)]) == 0',
        'count([p | c := input.review.object.spec.containers[_]; p := c.ports[_]; p.hostPort > input.parameters.n]) > 0',
        'count({v | v := input.review.object.metadata.labels[_]}) < count(input.review.object.metadata.labels)',
    ],
}
_RF_SRC = {"container": ["c := input.review.object.spec.containers[_]", "c := input_containers[_]", "some i\n  c := input.review.object.spec.containers[i]"],
           "label": ["v := input.review.object.metadata.labels[k]"], "volume": ["vol := input.review.object.spec.volumes[_]"], "none": [""]}
_RF_SUBJ = {"container": "c.name", "label": "k", "volume": "vol.name", "none": "input.review.object.metadata.name"}


def _rf_params(rnd):
    return {"prefix": rnd.choice(["gcr.io/", "openpolicyagent/", "/mnt", "ns-0", "", "registry.k8s.io/repo-1"]),
            "prefixes": rnd.sample(["gcr.io/", "quay.io/", "docker.io/library/", "openpolicyagent/", "registry.k8s.io/", "evil.example.com/repo-001"], rnd.randint(0, 3)),
            "cpu": rnd.choice(["100m", "2", "1", 4]), "mem": rnd.choice(["128Mi", "1Gi", "2G"]), "n": rnd.choice([0, 1, 2, 3, 8080, 1.5]),
            "name": rnd.choice(["c0", "c1", "vol-0", "pvc-1", "C0"]), "sub": rnd.choice(["repo-0", "latest", "v1", ""]),
            "tag": rnd.choice(["latest", "v1.0.18", "1.2.3"]), "tags": rnd.sample(["latest", "v1.0.18", "v2.1.0", "1.2.3"], rnd.randint(0, 3)),
            "pattern": rnd.choice(["^gcr[.]io/", "^(openpolicyagent|quay[.]io)/.+$", ":latest$", "^v[0-9]+$", "^team-[0-9]+$"]),
            "key": rnd.choice(["team", "label-06", "app", "label-15"]), "val": rnd.choice(["team-42", "v9", ""]),
            "labels": rnd.sample(["team", "label-01", "label-06", "label-22", "owner"], rnd.randint(0, 3)), "ns": rnd.choice(["ns-0001", "kube-system", "production"]),
            "flag": rnd.choice([True, False]), "volumes": rnd.sample(["emptyDir", "configMap", "secret", "hostPath", "persistentVolumeClaim", "projected"], rnd.randint(0, 4)),
            "pairs": [{"key": rnd.choice(["team", "label-06", "app", "label-15", "label-22"]), "val": rnd.choice(["team-42", "v9", "v123"]),
                       "pattern": rnd.choice(["^team-[0-9]+$", "^v[0-9]+$", "^x"])} for _ in range(rnd.randint(0, 3))],
            "mount": rnd.choice(["/mnt", "/", "/mnt/1"]), "policy": rnd.choice(["Always", "IfNotPresent"]), "kind": rnd.choice(["Pod", "Deployment"])}


def _rf_template(rnd, idx):
    bodies = []
    for b in range(rnd.choice([1, 1, 2, 3])):
        src = rnd.choice(["container", "container", "container", "label", "volume", "none", "none"])
        lines = [rnd.choice(_RF_SRC[src])] if _RF_SRC[src][0] else []
        conds = rnd.sample(_RF_CONDS[src], rnd.randint(1, 3))
        if src != "none" and rnd.random() < 0.3:
            conds.append(rnd.choice(_RF_CONDS["none"]))
        # two conditions of one body must not introduce the same local name twice
        seen, keep = set(), []
        for c in conds:
            intro = {ln.split(":=")[0].strip() for ln in c.split("\n") if ":=" in ln}
            if intro & seen:
                continue
            seen |= intro
            keep.append(c)
        rnd.shuffle(keep)
        lines += keep
        extra = rnd.choice(["input.parameters.n", "input.parameters.prefixes", "input.parameters.key", "input.review.object.metadata.name", "%d" % b])
        lines.append('msg := sprintf("b%d <%%v> <%%v>", [%s, %s])' % (b, _RF_SUBJ[src], extra))
        head = 'violation[{"msg": msg, "details": {"body": %d}}]' % b if rnd.random() < 0.3 else 'violation[{"msg": msg}]'
        bodies.append(head + " {\n  " + "\n  ".join(lines) + "\n}\n")
    return "package fz%d\nimport future.keywords.in\n" % idx + "".join(bodies) + _RF_HELPERS


def case_rego_fuzz(lib, n_templates=40, n_objects=120, seed=1, via_blob=False):
    """Random policies: `violation` bodies assembled from a menu of ~50 statement shapes of the in-tree templates (iteration
    over containers / labels / volumes, comprehensions, set difference, helper rules and multi-body functions, negation,
    builtins on object fields against parameters), each with three random parameter sets, against damaged synthetic Pods.
    Whatever the engine accepts must agree with the oracle result for result; rejected policies (rego_unsupported) are
    counted, never compared."""
    import random
    rnd = random.Random(seed)
    blob = W.synth_objects(52000 + seed, n_objects)
    revs = []
    for i in range(n_objects):
        o = json.loads(blob.get(i))
        o = _mutate(rnd, o, rnd.choice([0, 0, 0, 1, 2]))
        revs.append(D.Review(object=o))
    # single-container copies: a decision the device gets wrong for ONE container cannot hide behind another container's result
    # (a flagged pair the renderer finds nothing for is an engine error)
    for r in list(revs[:40]):
        cs = ((r.object.get("spec") or {}).get("containers") if isinstance(r.object.get("spec"), dict) else None) or []
        if isinstance(cs, list) and len(cs) > 1:
            for c in cs:
                o = json.loads(json.dumps(r.object))
                o["spec"]["containers"] = [c]
                revs.append(D.Review(object=o))
    # review shapes: UPDATE with an old object, DELETE (the old object is the one reviewed), user info, an explicit namespace object
    # (not through the blob path: a blob holds plain objects, which the ingest kernels flatten on the device)
    for i in range(0 if via_blob else len(revs)):
        r = rnd.random()
        o = revs[i].object
        if r < 0.12 and i:
            revs[i] = D.Review(object=o, old_object=revs[i - 1].object, operation="UPDATE", source="Original")
        elif r < 0.18:
            revs[i] = D.Review(object=None, old_object=o, operation="DELETE", source="Original")
        elif r < 0.30:
            revs[i] = D.Review(object=o, operation=rnd.choice(["CREATE", "UPDATE"]), source="Original",
                               user_info={"username": rnd.choice(["system:serviceaccount:kube-system:x", "alice", ""])})
        elif r < 0.40:
            revs[i] = D.Review(object=o, source="Original", namespace={"apiVersion": "v1", "kind": "Namespace", "metadata": dict(
                {"name": rnd.choice(["explicit-ns", "ns-0001"])}, **({"labels": {rnd.choice(["team", "label-06", "app"]): rnd.choice(["v9", "team-42"])}} if rnd.random() < 0.7 else {}))})
    n_objects = len(revs)
    accepted = n_results = n_device = 0
    rejected = []
    pyblob = W.PyBlob([r.object for r in revs]) if via_blob else None
    for t in range(n_templates):
        src = _rf_template(rnd, t)
        kind = "Fz%d" % t
        cons = [W._constraint(kind, "fz-%d-%d" % (t, i), params=_rf_params(rnd), action=rnd.choice([None, "warn"])) for i in range(3)]
        try:
            orc, drv, skipped = make_pair([(kind, src)], cons, lib_path=lib, skip_unsupported=True)
        except Exception as e:
            raise AssertionError("template %d of seed %d failed to load: %r\n%s" % (t, seed, e, src))
        if skipped:
            rejected.append((t, skipped[0][2]))
            if len(skipped) == 3 or skipped[0][1] is None:
                continue
        accepted += 1
        try:
            if via_blob:
                n_device += "ingest: device" in drv.Dump()
                resp = drv.ReviewBlob(pyblob, k8s.AUDIT_EP, flags=D.F_MATERIALIZE, source="")
            else:
                resp = drv.ReviewBatch(revs, k8s.AUDIT_EP)
        except D.GkError as e:
            m = __import__("re").search(r"object (\d+)\)", str(e))
            raise AssertionError("seed %d template %d: %s\n%s\nparams %s\nobject %s" % (
                seed, t, e, src, [c["spec"].get("parameters") for c in cons], json.dumps(revs[int(m.group(1))].object) if m else "?"))
        bad = {i for i, e in enumerate(resp.object_errors or []) if e}
        try:
            want = oracle_results_safe(orc, revs, k8s.AUDIT_EP, skip=bad)
            got = {x for x in engine_results(resp) if x[0] not in bad}
            assert_same(want, got)
        except AssertionError as e:
            raise AssertionError("seed %d template %d:\n%s\nparams %s\n%s" % (seed, t, src, [c["spec"].get("parameters") for c in cons], str(e)[:1500]))
        n_results += len(want)
        # the audit's lazy path (single-result proofs of the lowering + the ambiguity netlist) must count what the renderer renders
        if not bad:
            per = {}
            for r in resp.results:
                per[r.constraint] = per.get(r.constraint, 0) + 1
            run = D.AuditRun(drv, violations_limit=3)
            rb = drv.upload_blob(pyblob, source="") if via_blob else drv.upload(revs)
            run.add_batch(rb, k8s.AUDIT_EP)
            rep = run.report()
            try:
                assert {k: v for k, v in rep["totalViolations"].items() if v} == per
            except AssertionError:
                raise AssertionError("seed %d template %d: audit totals %s != rendered %s (counted %d, evaluated %d)\n%s\nparams %s" % (
                    seed, t, rep["totalViolations"], per, rep["pairsCounted"], rep["pairsEvaluated"], src, [c["spec"].get("parameters") for c in cons]))
            n_counted = locals().get("n_counted", 0) + rep["pairsCounted"]
    assert accepted >= n_templates * 0.6, rejected
    if via_blob:
        return accepted, n_results, rejected, n_device
    return accepted, n_results, rejected


def case_cross_scope_join(lib, n=3000):
    """Nested independent iterations tested together (containers x volumeMounts x volumes joined by name -- the shape of the
    host-filesystem policies when written without helper rules): lowered with the inner collection nested under the outer
    loop's rows; a conjunct that does not depend on the inner loop is hoisted out of it."""
    src = '''package j
violation[{"msg": msg}] {
  c := input.review.object.spec.containers[_]
  m := c.volumeMounts[_]
  vol := input.review.object.spec.volumes[_]
  vol.name == m.name
  vol.hostPath
  not m.readOnly
  startswith(vol.hostPath.path, input.parameters.prefix)
  msg := sprintf("container <%v> mounts hostPath volume <%v> read-write", [c.name, vol.name])
}
violation[{"msg": msg}] {
  input.review.object.spec.containers[_].name == input.review.object.spec.volumes[_].name
  msg := "a container is named like a volume"
}
violation[{"msg": msg}] {
  v := input.review.object.metadata.labels[k]
  input.review.object.spec.containers[_].securityContext.privileged
  startswith(k, input.parameters.label)
  msg := sprintf("privileged pod carries label <%v>=<%v>", [k, v])
}
'''
    blob = W.synth_objects(99, n)
    revs = [D.Review(object=json.loads(blob.get(i))) for i in range(n)]
    revs.append(D.Review(object={"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "same", "namespace": "default"},
                                 "spec": {"containers": [{"name": "x", "image": "i"}], "volumes": [{"name": "x", "emptyDir": {}}]}}))
    cons = [W._constraint("J", "j", params={"prefix": "/", "label": "label-0"}), W._constraint("J", "j2", params={"prefix": "/var", "label": "team"})]
    orc, drv, skipped = make_pair([("J", src)], cons, lib_path=lib)
    assert not skipped
    resp = drv.ReviewBatch(revs, k8s.AUDIT_EP)
    assert not any(resp.object_errors or [])
    want = oracle_results(orc, revs, k8s.AUDIT_EP)
    msgs = {w[2] for w in want}
    assert "a container is named like a volume" in msgs and any(m.startswith("container <") for m in msgs) and any(m.startswith("privileged pod") for m in msgs)
    assert_same(want, engine_results(resp))
    return len(want)


# ------------------------------------------------------------------------------------------ pkg/target vectors
DENY_ALL = 'package denyall\nviolation[{"msg": msg}] {\n  msg := "denyall constraint installed"\n}\n'   # target_integration_test.go:37-43


def _target_shapes(v):
    """The three request shapes TestConstraintEnforcement reviews (target_integration_test.go:446-520)."""
    ns = v["namespace"]
    nsn = ns["metadata"]["name"] if ns else ""
    yield "object", dict(object=v["object"], namespace=ns, namespace_name=nsn)
    yield "oldObject", dict(object=None, old_object=v["object"], namespace=ns, namespace_name=nsn)
    yield "unstructured", dict(object=v["object"], namespace=ns)


def case_target_enforcement(lib):
    """26 scenarios the reference runs through its real client + Rego driver: allowed <=> no results."""
    n = 0
    for v in golden("target_vectors.json")["constraint_enforcement"]:
        orc, drv, _ = make_pair([("DenyAll", DENY_ALL)], [v["constraint"]], lib_path=lib)
        for shape, kw in _target_shapes(v):
            rev = D.Review(**kw)
            resp = drv.ReviewBatch([rev], k8s.AUDIT_EP)
            assert (len(resp.results) == 0) == v["allowed"], (v["name"], shape, [r.msg for r in resp.results])
            assert_same(oracle_results(orc, [rev], k8s.AUDIT_EP), engine_results(resp))
            n += 1
    assert n == 78


def _matcher_review(req):
    kw = dict(object=req.get("object"), old_object=req.get("oldObject"), namespace=req.get("namespace"))
    if req["shape"] in ("AugmentedReview", "AdmissionRequest"):
        kw["namespace_name"] = req.get("namespaceName", "")     # AdmissionRequest.Namespace is explicit (matcher.go:37-39)
    return kw


def case_target_matcher(lib):
    """TestMatcher_Match: object OR old object, review namespace vs cached namespace, the two error classes."""
    t = golden("templates.json")["fixtures_TemplateNeverValidate"]
    n = 0
    for v in golden("target_vectors.json")["matcher_match"]:
        if v["request"] is None or v["match"] is None:
            continue                                               # "nil": HandleReview does not handle a nil request
        con = {"kind": t["kind"], "metadata": {"name": "c"}, "spec": {"match": v["match"]}}
        nss = [v["cachedNamespace"]] if v["cachedNamespace"] else []
        orc, drv, _ = make_pair([(t["kind"], t["rego"])], [con], nss, lib_path=lib)
        rev = D.Review(**_matcher_review(v["request"]))
        resp = drv.ReviewBatch([rev], k8s.AUDIT_EP)
        obj_err = (resp.object_errors or [None])[0]
        flagged = bool(resp.viol_bits[0, 0] & 1) and not obj_err
        errored = bool(resp.err_bits[0, 0] & 1)
        if v["wantErr"] == "ErrRequestObject":
            assert obj_err or (resp.results and resp.results[0].autoreject and "invalid request object" in resp.results[0].msg), v["name"]
            assert not flagged
        elif v["wantErr"] == "ErrMatching":
            assert errored and resp.results[0].autoreject and resp.results[0].msg.startswith("unable to match constraints: error matching the requested object"), v["name"]
        else:
            assert not errored and not obj_err, (v["name"], obj_err)
            assert flagged == v["want"], v["name"]
        if not obj_err:
            assert_same(oracle_results(orc, [rev], k8s.AUDIT_EP), engine_results(resp))
        n += 1
    assert n >= 15


# ------------------------------------------------------------------------------------------ gator TestTest table
def case_gator_test_table(lib):
    """pkg/gator/test/test_test.go:85-268: every input document is reviewed at the gator enforcement point; the exact
    list of (message, constraint, action, scoped actions) is what the reference asserts.  gator adds every input object
    as data before auditing (pkg/gator/test/test.go:91-97), which is what the referential rows (data.inventory) read."""
    import pytest
    n = 0
    for row in golden("gator_test_table.json"):
        tm, cons, nss = _split_docs(row["docs"])
        if row["wantErr"]:
            drv = D.Driver(lib_path=lib)
            with pytest.raises(D.GkError, match="no template|template"):
                for c in cons:
                    drv.AddConstraint(c)
            continue
        orc, drv, _ = make_pair(tm, cons, nss, lib_path=lib)
        for d in row["docs"]:
            if d.get("kind") not in ("ConstraintTemplate", "Namespace") and not str(d.get("apiVersion", "")).startswith("constraints.gatekeeper.sh"):
                orc.add_data(d)
                drv.AddData("admission.k8s.gatekeeper.sh", orc.data_path(d), d)
        revs = [D.Review(object=d) for d in row["docs"]]
        resp = drv.ReviewBatch(revs, k8s.GATOR_EP)
        assert_same(oracle_results(orc, revs, k8s.GATOR_EP), engine_results(resp))
        got = sorted((r.msg, r.constraint) + ((r.enforcement_action, tuple(r.scoped_enforcement_actions)) if any("enforcementAction" in w for w in row["want"]) else ())
                     for r in resp.results)
        want = sorted((w["msg"], "%s/%s" % (w["constraintKind"], w["constraint"])) +
                      ((w["enforcementAction"], tuple(w.get("scopedEnforcementActions", []))) if "enforcementAction" in w else ())
                      for w in row["want"])
        assert got == want, (row["name"], got, want)
        n += 1
    assert n >= 6


def case_verify_suite(lib):
    """test/gator/verify/suite.yaml:1-37 through the engine (K8sFooIs: object.get(input, "parameters", {})): the
    allow / deny expectation of every non-expansion case, and parity with the oracle."""
    vs = golden("verify_suite.json")
    files = vs["files"]
    tmpl = k8s.template_from_yaml_obj(vs["template"])
    expect = [("constraint.yaml", "allow_foo.yaml", False), ("constraint.yaml", "deny_foo.yaml", True),
              ("constraint_with_scopedEA.yaml", "allow_foo.yaml", False), ("constraint_with_scopedEA.yaml", "deny_foo.yaml", True),
              ("constraint_with_scopedEA_without_gator_ep.yaml", "allow_foo.yaml", False),
              ("constraint_with_scopedEA_without_gator_ep.yaml", "deny_foo.yaml", False)]
    for cfile, ofile, violations in expect:
        orc, drv, skipped = make_pair([tmpl], [files[cfile][0]], lib_path=lib)
        assert not skipped
        revs = [D.Review(object=files[ofile][0])]
        resp = drv.ReviewBatch(revs, k8s.GATOR_EP)
        assert bool(resp.results) == violations, (cfile, ofile, [r.msg for r in resp.results])
        assert_same(oracle_results(orc, revs, k8s.GATOR_EP), engine_results(resp))


def case_inexact_numbers(lib):
    """Numbers that are not exact int64 (fractions, 1e30, -1e25) must be COMPARED, not make the object skip every
    constraint (round-1 advisor finding: `spec.replicas: 5.5` returned no results at all)."""
    rego = """package k8smaxreplicas
violation[{"msg": msg}] {
  r := input.review.object.spec.replicas
  r > input.parameters.max
  msg := sprintf("replicas %v > %v", [r, input.parameters.max])
}
violation[{"msg": msg}] {
  r := input.review.object.spec.replicas
  r < input.parameters.min
  msg := sprintf("replicas %v < %v", [r, input.parameters.min])
}
violation[{"msg": "exactly five"}] { input.review.object.spec.replicas == 5 }
violation[{"msg": "not five"}] { input.review.object.spec.replicas != 5 }
"""
    owner = """package k8snoowner
violation[{"msg": "no owner"}] { not input.review.object.metadata.labels.owner }
"""
    tm = [("K8sMaxReplicas", rego), ("K8sNoOwner", owner)]
    cons = [{"apiVersion": "constraints.gatekeeper.sh/v1beta1", "kind": "K8sMaxReplicas", "metadata": {"name": "r%d" % i},
             "spec": {"parameters": {"max": mx, "min": mn}}} for i, (mn, mx) in enumerate([(1, 3), (5, 5), (-2, 6), (0, 2.5), (5.5, 7)])]
    cons.append({"apiVersion": "constraints.gatekeeper.sh/v1beta1", "kind": "K8sNoOwner", "metadata": {"name": "owner"}, "spec": {}})
    orc, drv, _ = make_pair(tm, cons, lib_path=lib)
    vals = [5.5, 5, 5.0, 3, 3.0001, 2.9999, 1e30, -1e25, -0.5, 0, 2, 2.5, 6, 6.5, 9223372036854775807, 9223372036854775808,
            -9223372036854775809, 1.5e3, 0.1, "5", None, True, [5], {"a": 5}]
    revs = [D.Review(object={"apiVersion": "apps/v1", "kind": "Deployment", "metadata": {"name": "d%d" % i, "namespace": "default"},
                             "spec": {"replicas": v}}) for i, v in enumerate(vals)]
    resp = drv.ReviewBatch(revs, k8s.AUDIT_EP)
    assert not any(resp.object_errors), resp.object_errors
    want = oracle_results(orc, revs, k8s.AUDIT_EP)
    # documented limit: a fractional value against a fractional threshold with the same integer part (2.5 vs max 2.5 is exact
    # equality here and is fine; 5.5 vs min 5.5 too) -- the vectors above avoid the ambiguous "5.7 vs 5.5" shape
    assert_same(want, engine_results(resp))
    assert sum(1 for w in want if w[1] == "K8sNoOwner/owner") == len(vals)
    return resp


# ------------------------------------------------------------------------------------------ device ingest (blob path)
def _blob_parity(orc, drv, docs, ep=k8s.AUDIT_EP, source="Original", expect_errors=()):
    """Raw JSON documents through gk_review_blob (flattened by the ingest kernels when the snapshot allows it) against the oracle
    on the parsed documents; `docs` are bytes or objects."""
    blob = W.PyBlob(docs)
    resp = drv.ReviewBlob(blob, ep, flags=D.F_MATERIALIZE, source=source)
    errs = resp.object_errors or [None] * len(docs)
    bad = {i for i, e in enumerate(errs) if e}
    assert bad == set(expect_errors), (sorted(bad), [errs[i] for i in sorted(bad)][:5])
    revs = []
    for i, d in enumerate(docs):
        if i in bad:
            revs.append(None)
            continue
        revs.append(D.Review(object=json.loads(d) if isinstance(d, (bytes, bytearray)) else d, source=source))
    want = set()
    for i, r in enumerate(revs):
        if r is None:
            continue
        for x in oracle_results(orc, [r], ep):
            want.add((i,) + x[1:])
    assert_same(want, {x for x in engine_results(resp) if x[0] not in bad})
    return resp, want


def case_blob_config2(lib, n=1500, start=0):
    tm, cons = W.config2()
    orc, drv, _ = make_pair(tm, cons, W.synth_namespaces(), lib_path=lib)
    assert "ingest: device" in drv.Dump(), drv.Dump().splitlines()[1]
    blob = W.synth_objects(start, n)
    docs = [blob.get(i) for i in range(n)]
    resp, want = _blob_parity(orc, drv, docs)
    # the same page through the host flattener gives the same bitmap
    host = drv.ReviewBatch([D.Review(object=json.loads(d), source="Original") for d in docs], k8s.AUDIT_EP, materialize=False)
    assert (host.viol_bits == resp.viol_bits).all() and host.totals == resp.totals
    return resp, want


def case_blob_fuzz(lib, n=600, seed=4321, start=9000):
    """config-2 constraints against structurally damaged Pods, raw JSON in, device-flattened."""
    import random
    rnd = random.Random(seed)
    tm, cons = W.config2()
    orc, drv, _ = make_pair(tm, cons, W.synth_namespaces(), lib_path=lib)
    blob = W.synth_objects(start, n)
    docs = [_mutate(rnd, json.loads(blob.get(i)), rnd.choice([0, 1, 1, 2, 3, 5])) for i in range(n)]
    return _blob_parity(orc, drv, docs)


def case_blob_json_oddities(lib):
    """What a JSON encoder may legally (or illegally) write: whitespace, escapes in keys and values, surrogate pairs, duplicate
    keys (the last one wins), exponents / fractions / huge numbers, empty containers, and documents the review must refuse."""
    tm, cons = W.config2()
    orc, drv, _ = make_pair(tm, cons, W.synth_namespaces(), lib_path=lib)
    pod = json.loads(W.synth_objects(77, 1).get(0))
    pod["spec"]["containers"][0]["image"] = "registry.k8s.io/a\"b\\c:la\ttest"
    pod["metadata"]["labels"] = {"te\u0061m": "x", "team": "pl\u00e4tform", "emoji": "\ud83d\ude00"}
    good = json.dumps(pod)
    docs = [
        good.encode(),
        json.dumps(pod, indent=2).encode(),                                # whitespace everywhere
        good.replace('"image"', '"\\u0069mage"').encode(),                 # an escaped key
        good.replace('"spec":', '"spec": {"containers": "shadowed"}, "spec":').encode(),   # duplicate key: the last wins
        good.replace('"kind":"Pod"', '"kind":"Job","kind":"Pod"').encode(),
        json.dumps(dict(pod, spec=dict(pod["spec"], hostNetwork=True, priority=1e3, ratio=1.50, big=123456789012345678901234567890,
                                       neg=-0.0, tiny=1e-7, exp=12E+2))).encode(),
        json.dumps(dict(pod, spec={})).encode(), json.dumps(dict(pod, spec=[])).encode(), json.dumps(dict(pod, metadata=None)).encode(),
        b'{"apiVersion":"v1","kind":"Pod","metadata":{"name":"x","namespace":"default","labels":{}},"spec":{"containers":[]}}',
        b'  {"apiVersion" : "apps/v1" , "kind" : "Deployment", "metadata" : {"name":"d"} , "spec" : {"replicas" : 3.0e0 } }  ',
        # ---- refused at the review boundary
        b'{"apiVersion":"v1","kind":"Pod","metadata":{"name":"x"}',         # truncated
        b'{"apiVersion":"v1","kind":"Pod",}',                               # trailing comma
        b'{"apiVersion":"v1","kind":"Pod"} x',                              # trailing characters
        b'{"apiVersion":"v1","metadata":{"name":"nokind"}}',                # kind missing
        b'{"apiVersion":"v1","kind":"","metadata":{}}',                     # kind empty
        b'[1,2,3]', b'"a string"', b'{"kind":"Pod","a":01e}',
        b'{"kind":"Pod","a":"bad \\q escape"}', b'{"kind":"Pod","a":1.2.3}', b'{"kind":"Pod","a":-}',
        b'{"kind":"Pod","apiVersion":"v1","metadata":{"name":"ok-after-errors","labels":{"team":"a"}},"spec":{"containers":[{"name":"c","image":"x:latest"}]}}',
    ]
    # the tokeniser scans strings a 64-bit word at a time: closing quotes, escapes and unterminated strings at every byte alignment
    sweep_ok, sweep_bad = [], []
    for pad in range(9):
        for ln in (0, 1, 6, 7, 8, 9, 15, 16, 17, 31):
            for esc_at in (None, 0, ln // 2, max(ln - 1, 0)):
                body = "x" * ln
                if esc_at is not None and ln:
                    body = body[:esc_at] + (chr(92) * 2 if (pad + ln) % 2 else chr(92) + chr(34)) + body[esc_at + 1:]
                doc = '{"apiVersion":"v1","kind":"Pod","metadata":{"name":"p%s","namespace":"default","labels":{"team":"%s"}},"spec":{"containers":[{"name":"c","image":"gcr.io/%s:1"}]}}' % (
                    "y" * pad, body, body)
                sweep_ok.append(doc.encode())
        sweep_bad.append(('{"apiVersion":"v1","kind":"Pod","metadata":{"name":"%s","labels":{"team":"ab' % ("y" * pad) + chr(92)).encode())               # ends inside an escape
        sweep_bad.append(('{"apiVersion":"v1","kind":"Pod","metadata":{"name":"%s","labels":{"team":"ab' % ("y" * pad) + chr(92) + 'u12"}}}').encode())    # short \\u escape
    docs = docs[:-1] + sweep_bad + [docs[-1]] + sweep_ok
    n_tail = 1 + len(sweep_ok)
    first_bad = next(i for i, d in enumerate(docs) if d.startswith(b'{"apiVersion":"v1","kind":"Pod","metadata":{"name":"x"}') and not d.endswith(b"}}"))
    bad = set(range(first_bad, len(docs) - n_tail))
    resp, want = _blob_parity(orc, drv, docs, expect_errors=bad)
    # every refusal carries the host parser's wording
    host = drv.ReviewBatch([D.Review(object=d, source="Original") for d in docs], k8s.AUDIT_EP)
    assert [bool(e) for e in host.object_errors] == [bool(e) for e in resp.object_errors]
    assert [e for e in host.object_errors if e] == [e for e in resp.object_errors if e]
    return resp, want


def case_blob_other_templates(lib, n=300, seed=7):
    """The in-tree templates outside config 2 (regex labels, object.get defaults, custom fields, comprehension allowed repos ...),
    one engine per template, raw JSON in: device-flattened where the snapshot allows it, host-flattened otherwise."""
    import random
    rnd = random.Random(seed)
    t = golden("templates.json")
    pick = {"requiredlabels_agilebank": [{"labels": [{"key": "owner", "allowedRegex": "^[a-z]+[.]agilebank[.]demo$"}, {"key": "team"}]},
                                         {"message": "custom message", "labels": [{"key": "team", "allowedRegex": "^(a|b|team-[0-9]+)$"}]}],
            "requiredlabels_regov1": [{"labels": ["team", "owner"]}, {"labels": []}],
            "fooischeck": [{"foo": "bar"}, {"foo": ""}, {}, {"foo": 7}],
            "namespacelabelcheck": [{"requiredLabel": "team"}],
            "allowedrepos": [{"repos": ["gcr.io/", "quay.io/"]}, {"repos": []}, {"repos": ["openpolicyagent/opa:", "docker.io/library/nginx"]}],
            "fixtures_TemplateRestrictCustomField": [{"expectedCustomField": "x"}, {"expectedCustomField": 7}, {"expectedCustomField": {"a": [1, 2]}}]}
    blob = W.synth_objects(30000, n)
    docs = []
    for i in range(n):
        o = json.loads(blob.get(i))
        if rnd.random() < 0.3:
            o.setdefault("metadata", {}).setdefault("labels", {})[rnd.choice(["owner", "team", "app"])] = rnd.choice(
                ["alice.agilebank.demo", "bob", "team-7", "a", "", "Team-1", "x.agilebank.demo.evil"])
        if rnd.random() < 0.3:
            o["foo"] = rnd.choice(["bar", "", "baz", 7, None, ["bar"]])
        if rnd.random() < 0.3:
            o.setdefault("spec", {})["customField"] = rnd.choice(["x", "y", 7, 7.0, {"a": [1, 2]}, {"a": [2, 1]}, None, [1]])
        docs.append(_mutate(rnd, o, rnd.choice([0, 0, 1, 2])))
    n_dev = n_host = 0
    for name, plist in pick.items():
        cons = [W._constraint(t[name]["kind"], "%s-%d" % (name.replace("_", "-").lower(), i), params=params) for i, params in enumerate(plist)]
        orc, drv, skipped = make_pair([(t[name]["kind"], t[name]["rego"])], cons, W.synth_namespaces(), lib_path=lib, skip_unsupported=True)
        assert not skipped, skipped
        if "ingest: device" in drv.Dump():
            n_dev += 1
        else:
            n_host += 1
        _blob_parity(orc, drv, docs)
    assert n_dev >= 4, (n_dev, n_host)
    return n_dev, n_host


# ------------------------------------------------------------------------------------------ reference-held pins outside the test tables
def case_doc_pins(lib):
    """The exact messages the reference's documentation quotes (website/docs/violations.md, workload-resources.md, audit.md) and
    the AllowedRepos manifests of the gator suite (tests/golden/make_doc_pins.py): oracle AND engine must print them verbatim --
    they pin sprintf("%v") of an array, of an object and of strings for K8sAllowedRepos, K8sPSPPrivilegedContainer, K8sContainerLimits."""
    n = 0
    for case in golden("doc_pins.json"):
        tm, cons, nss = _split_docs(case["docs"])
        orc, drv, skipped = make_pair(tm, cons, nss, lib_path=lib)
        assert not skipped
        objs = [d for d in case["docs"] if d.get("kind") not in ("ConstraintTemplate",) and not str(d.get("apiVersion", "")).startswith("constraints.gatekeeper.sh")]
        revs = [D.Review(object=d, source="Original") for d in objs]
        ep = case["ep"]
        want = oracle_results(orc, revs, ep)
        for via_blob in (False, True):
            resp = drv.ReviewBlob(W.PyBlob(objs), ep, flags=D.F_MATERIALIZE) if via_blob else drv.ReviewBatch(revs, ep)
            got = engine_results(resp)
            assert_same(want, got)
            have = sorted((r[1], r[2], r[4]) for r in got)
            exp = sorted((e["constraint"], e["msg"], e["action"]) for e in case["expect"])
            assert have == exp, (case["name"], have, exp)
            assert bool(got) == bool(case.get("expect_any", bool(exp)))
        n += 1
    return n


def case_wildcard_vectors_through_kernel(lib):
    """pkg/wildcard/wildcard_test.go:7-193 through the DEVICE matcher (vm_core.h gk_wild / gk_wild_gen): `matches` vectors as
    spec.match.namespaces and as spec.match.name, `generateName` vectors as spec.match.name against metadata.generateName."""
    t = golden("templates.json")["fixtures_TemplateNeverValidate"]
    checked = 0
    for v in golden("wildcard_vectors.json"):
        pats = []
        if v["fn"] == "matches":
            pats.append(({"namespaces": [v["w"]]}, {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "p", "namespace": v["candidate"]}}))
            pats.append(({"excludedNamespaces": [v["w"]]}, {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "p", "namespace": v["candidate"]}}))
            pats.append(({"name": v["w"]}, {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": v["candidate"], "namespace": "x"}}))
        else:
            pats.append(({"name": v["w"]}, {"apiVersion": "v1", "kind": "Pod", "metadata": {"generateName": v["candidate"], "namespace": "x"}}))
        for match, obj in pats:
            if not obj["metadata"].get("namespace", "x") or obj["metadata"].get("name") == "":
                continue                                  # an empty namespace / name is "not set": another branch of match.go
            drv = D.Driver(lib_path=lib)
            drv.add_template(t["kind"], t["rego"])
            drv.AddConstraint({"kind": t["kind"], "metadata": {"name": "c"}, "spec": {"match": match}})
            want = v["matches"] != ("excludedNamespaces" in match)
            for via_blob in (False, True):
                resp = (drv.ReviewBlob(W.PyBlob([obj]), k8s.AUDIT_EP, source="") if via_blob else drv.ReviewBatch([D.Review(object=obj)], k8s.AUDIT_EP))
                assert bool(resp.viol_bits[0, 0] & 1) == want, (v, match, via_blob)
            checked += 1
    assert checked >= 40
    return checked


# ------------------------------------------------------------------------------------------ expansion (a-16 / f-3)
class pytest_raises_gk:
    """`with pytest_raises_gk(D, text):` -- the block must raise D.GkError whose message contains `text`"""
    def __init__(self, D_, text):
        self.D, self.text = D_, text

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        assert et is not None and issubclass(et, self.D.GkError), "expected GkError containing %r" % self.text
        assert self.text in str(ev), (self.text, str(ev))
        return True


def case_expansion(lib):
    """ExpansionTemplates through the review batch: (1) the reference's TestExpand vectors (resultant objects) on the oracle,
    (2) the gator expansion manifests with the messages test.bats asserts, on oracle and engine, (3) a mixed page: Deployments,
    Pods and a nested generator chain, with and without an action override, engine == oracle result for result."""
    from oracle import expansion as X
    vec = golden("expansion_vectors.json")
    for c in vec["expand"]:
        s = X.System()
        for t in c["templates"]:
            s.upsert(t)
        try:
            got, err = s.expand(c["generator"], c["ns"]), False
        except X.ExpansionError:
            got, err = [], True
        assert err == c["expectErr"], c["name"]
        if not err:
            assert sorted(json.dumps(list(x), sort_keys=True) for x in got) == sorted(
                json.dumps([w["obj"], w["templateName"], w["enforcementAction"]], sort_keys=True) for w in c["want"]), c["name"]

    # (1b) TestValidateTemplate (system_test.go:311-424): the substring each refusal must contain -- oracle and engine
    drv0 = D.Driver(lib_path=lib)
    for c in vec["validate"]:
        errs = []
        try:
            X.validate_template(c["template"])
            errs.append(None)
        except X.ExpansionError as e:
            errs.append(str(e))
        try:
            drv0.AddExpansionTemplate(c["template"])
            errs.append(None)
            drv0.RemoveExpansionTemplate(c["template"]["metadata"]["name"])
        except D.GkError as e:
            errs.append(str(e))
        for e in errs:
            assert (e is None) if c["errSubstr"] is None else (e is not None and c["errSubstr"] in e), (c["name"], errs)
        if errs[0] is not None:
            assert errs[0] in errs[1], (c["name"], errs)     # the same text on both sides
    drv0.close()

    # (1b') TestDB (db_test.go:27-647): upserts / removals -> which stored templates are set aside as part of an expansion cycle
    # (GetConflicts), and which upsert reports "template forms expansion cycle" while storing the template all the same
    for c in vec["db"]:
        xs, drv2 = X.System(), D.Driver(lib_path=lib)
        for op in c["ops"]:
            name = op["template"]["metadata"]["name"]
            if op["op"] == "remove":
                xs.remove(name)
                drv2.RemoveExpansionTemplate(name)
                continue
            errs = []
            try:
                xs.upsert(op["template"])
                errs.append(None)
            except X.ExpansionError as e:
                errs.append(str(e))
            try:
                drv2.AddExpansionTemplate(op["template"])
                errs.append(None)
            except D.GkError as e:
                errs.append(str(e))
            for e in errs:
                assert (e is not None and "template forms expansion cycle" in e) if op["wantErr"] else e is None, (c["name"], name, errs)
        want = sorted(n for n, bad in c["want"].items() if bad)
        assert sorted(xs.conflicts()) == want, (c["name"], sorted(xs.conflicts()), want)
        assert drv2.ExpansionConflicts() == want, (c["name"], drv2.ExpansionConflicts(), want)
        assert sorted(xs.templates) == sorted(c["want"]), (c["name"], sorted(xs.templates))     # the store: cyclic templates are kept
        drv2.close()

    # (1c) TestExpandResource (system_test.go:426-660): expandResource() itself on the oracle; through a review on the engine, where a
    # constraint that copies the reviewed object into its details shows the resultant (the vectors' templates carry no applyTo --
    # expandResource does not look at it -- so the engine path adds the parent's GVK; the two vectors whose template is not a valid
    # template at all cannot be registered and stay oracle-only, the engine must refuse them)
    dump_kind = "K8sDumpObject"
    dump_rego = ('package k8sdumpobject\n\nviolation[{"msg": msg, "details": {"obj": input.review.object}}] {\n'
                 '  msg := sprintf("%v", [input.review.object.metadata.name])\n}\n')
    for c in vec["expand_resource"]:
        try:
            got, err = X.expand_resource(c["obj"], c["ns"], c["template"]), None
        except X.ExpansionError as e:
            got, err = None, str(e)
        if c["errSubstr"] is not None:
            assert err is not None and c["errSubstr"] in err, (c["name"], err)
        else:
            assert err is None and (c["want"] is None or got == c["want"]), (c["name"], err, got)
        tdoc = json.loads(json.dumps(c["template"]))
        g, v, k = X._gvk(c["obj"])
        tdoc["spec"]["applyTo"] = [{"groups": [g], "versions": [v], "kinds": [k]}]
        drv1 = D.Driver(lib_path=lib)
        drv1.add_template(dump_kind, dump_rego)
        drv1.AddConstraint(W._constraint(dump_kind, "dump", match={"kinds": [{"apiGroups": [""], "kinds": ["Pod"]}]}))
        try:
            X.validate_template(tdoc)
        except X.ExpansionError as e:
            with pytest_raises_gk(D, str(e)):
                drv1.AddExpansionTemplate(tdoc)
            drv1.close()
            continue
        drv1.AddExpansionTemplate(tdoc)
        nsobj = {"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": c["ns"]}} if c["ns"] is not None else None
        resp = drv1.ReviewBatch([D.Review(object=c["obj"], namespace=nsobj, source="Original")], k8s.AUDIT_EP)
        if err is not None:
            assert resp.object_errors and resp.object_errors[0] == "unable to expand object: " + err, (c["name"], resp.object_errors, err)
            assert c["errSubstr"] in resp.object_errors[0]
        else:
            assert not (resp.object_errors and resp.object_errors[0]), (c["name"], resp.object_errors)
            assert len(resp.results) == 1, (c["name"], [r.msg for r in resp.results])
            r = resp.results[0]
            assert r.msg == "[Implied by %s] %s" % (tdoc["metadata"]["name"], got["metadata"]["name"]), r.msg
            assert r.details == {"obj": got}, (c["name"], r.details, got)
        drv1.close()

    # (1c') the webhook's generator: obj.SetNamespace(req.Namespace) before expansion (pkg/webhook/policy.go:608) -- with no Namespace
    # object for the review the resultant takes the request's namespace, not the one written in the object
    c = vec["expand_resource"][1]     # "successful expansion without namespace": the Deployment carries its own metadata.namespace
    tdoc = json.loads(json.dumps(c["template"]))
    g, v, k = X._gvk(c["obj"])
    tdoc["spec"]["applyTo"] = [{"groups": [g], "versions": [v], "kinds": [k]}]
    drv4 = D.Driver(lib_path=lib)
    drv4.add_template(dump_kind, dump_rego)
    drv4.AddConstraint(W._constraint(dump_kind, "dump", match={"kinds": [{"apiGroups": [""], "kinds": ["Pod"]}]}))
    drv4.AddExpansionTemplate(tdoc)
    for req_ns, want_ns in (("elsewhere", "elsewhere"), ("", None), (None, c["want"]["metadata"]["namespace"])):
        resp = drv4.ReviewBatch([D.Review(object=c["obj"], operation="CREATE", namespace_name=req_ns, source="Original")], k8s.WEBHOOK_EP)
        assert len(resp.results) == 1, (req_ns, [r.msg for r in resp.results], resp.object_errors)
        assert resp.results[0].details["obj"]["metadata"].get("namespace") == want_ns, (req_ns, resp.results[0].details["obj"]["metadata"])
    drv4.close()

    # (1d) TestApplyTo (pkg/mutation/match/match_test.go:717-845): ApplyTo.Matches decides which templates expand a GVK -- on the oracle's
    # relation, and on the engine as "a generator of that GVK has / has no resultant"
    drv3 = D.Driver(lib_path=lib)
    drv3.add_template(dump_kind, dump_rego)
    drv3.AddConstraint(W._constraint(dump_kind, "dump", match={"kinds": [{"apiGroups": [""], "kinds": ["Pod"]}]}))
    for c in vec["apply_to"]:
        g, v, k = c["gvk"]
        assert any(X._apply_matches(a, (g, v, k)) for a in c["applyTo"]) == c["wantApply"], c["name"]
        tdoc = {"apiVersion": "expansion.gatekeeper.sh/v1beta1", "kind": "ExpansionTemplate", "metadata": {"name": "apply-to"},
                "spec": {"applyTo": c["applyTo"], "templateSource": "spec.template", "generatedGVK": {"group": "", "version": "v1", "kind": "Pod"}}}
        drv3.AddExpansionTemplate(tdoc)
        gen = {"apiVersion": (g + "/" + v) if g else v, "kind": k, "metadata": {"name": "gen"}, "spec": {"template": {"metadata": {"labels": {"a": "b"}}}}}
        resp = drv3.ReviewBatch([D.Review(object=gen, source="Original")], k8s.AUDIT_EP)
        assert [r.msg for r in resp.results] == (["[Implied by apply-to] gen-pod"] if c["wantApply"] else []), (c["name"], [r.msg for r in resp.results])
        drv3.RemoveExpansionTemplate("apply-to")
    drv3.close()

    def run(docs, ep, drv_revs=None):
        tm, cons, nss = _split_docs(docs)
        orc, drv, skipped = make_pair(tm, cons, nss, lib_path=lib)
        assert not skipped
        xs = X.System()
        for d in docs:
            if d.get("kind") == "ExpansionTemplate":
                xs.upsert(d)
                drv.AddExpansionTemplate(d)
        objs = [d for d in docs if d.get("kind") not in ("ConstraintTemplate", "ExpansionTemplate") and not str(d.get("apiVersion", "")).startswith("constraints.gatekeeper.sh")]
        revs = [D.Review(object=d, source="Original") for d in objs]
        want = set()
        for i, r in enumerate(revs):
            rv = k8s.Review(obj=r.object, source="Original")
            for x in X.review_with_expansion(orc, xs, rv, ep):
                want.add((i, "%s/%s" % x["constraint"], x["msg"], json.dumps(x["details"], sort_keys=True), x["enforcementAction"],
                          tuple(x["scopedEnforcementActions"]), bool(x.get("autoreject"))))
        resp = drv.ReviewBatch(revs, ep)
        assert_same(want, engine_results(resp))
        return resp, want

    g = vec["gator"]
    resp, want = run(g["docs"], k8s.GATOR_EP)
    assert any(g["without_ns_substring"] in w[2] for w in want), [w[2] for w in want]
    resp, want = run(g["docs"] + g["ns_docs"], k8s.GATOR_EP)
    assert any(g["with_ns_substring"] in w[2] for w in want), [w[2] for w in want]

    # (3) a mixed page
    t = golden("templates.json")
    tmpl = lambda name, kinds, gen, src="spec.template", action=None, groups=("apps",): {
        "apiVersion": "expansion.gatekeeper.sh/v1alpha1", "kind": "ExpansionTemplate", "metadata": {"name": name},
        "spec": dict({"applyTo": [{"groups": list(groups), "versions": ["v1"], "kinds": kinds}], "templateSource": src,
                      "generatedGVK": {"group": gen[0], "version": gen[1], "kind": gen[2]}}, **({"enforcementAction": action} if action else {}))}
    podspec = {"metadata": {"labels": {"app": "x"}}, "spec": {"containers": [{"name": "c", "image": "evil.example.com/a:latest",
                                                                               "securityContext": {"privileged": True}}]}}
    docs = [
        {"apiVersion": "templates.gatekeeper.sh/v1", "kind": "ConstraintTemplate", "metadata": {"name": "k8sallowedrepos"},
         "spec": {"crd": {"spec": {"names": {"kind": t["allowedrepos"]["kind"]}}}, "targets": [{"target": "admission.k8s.gatekeeper.sh", "rego": t["allowedrepos"]["rego"]}]}},
        {"apiVersion": "templates.gatekeeper.sh/v1", "kind": "ConstraintTemplate", "metadata": {"name": "psp"},
         "spec": {"crd": {"spec": {"names": {"kind": t["psp_privileged"]["kind"]}}}, "targets": [{"target": "admission.k8s.gatekeeper.sh", "rego": t["psp_privileged"]["rego"]}]}},
        W._constraint(t["allowedrepos"]["kind"], "repos", match=dict(W.POD), params={"repos": ["gcr.io/"]}, action="warn"),
        W._constraint(t["psp_privileged"]["kind"], "priv", match={"kinds": [{"apiGroups": [""], "kinds": ["Pod"]}], "source": "Generated"}),
        # a scoped constraint under a template that overrides the action: OverrideEnforcementAction (aggregate.go:47-58) replaces
        # EnforcementAction only, the result keeps its ScopedEnforcementActions
        W._constraint(t["psp_privileged"]["kind"], "priv-scoped", match=dict(W.POD), action="scoped",
                      scoped=[{"action": "warn", "enforcementPoints": [{"name": k8s.AUDIT_EP}]}, {"action": "deny", "enforcementPoints": [{"name": k8s.WEBHOOK_EP}]}]),
        tmpl("expand-deployments", ["Deployment", "ReplicaSet"], ("", "v1", "Pod")),
        tmpl("expand-cronjobs", ["CronJob"], ("batch", "v1", "Job"), src="spec.jobTemplate", groups=("batch",)),
        tmpl("expand-jobs", ["Job"], ("", "v1", "Pod"), action="dryrun", groups=("batch",)),
        {"apiVersion": "apps/v1", "kind": "Deployment", "metadata": {"name": "Web", "namespace": "prod"}, "spec": {"replicas": 2, "template": podspec}},
        {"apiVersion": "apps/v1", "kind": "ReplicaSet", "metadata": {"name": "rs"}, "spec": {"template": podspec}},
        {"apiVersion": "batch/v1", "kind": "CronJob", "metadata": {"name": "nightly", "namespace": "ops"}, "spec": {"jobTemplate": {"spec": {"template": podspec}}}},
        dict(podspec, apiVersion="v1", kind="Pod", metadata={"name": "plain", "namespace": "prod"}),
        {"apiVersion": "apps/v1", "kind": "Deployment", "metadata": {"name": "no-template"}, "spec": {"replicas": 1}},
        {"apiVersion": "v1", "kind": "ConfigMap", "metadata": {"name": "cm"}, "data": {}},
    ]
    for ep in (k8s.AUDIT_EP, k8s.WEBHOOK_EP):
        tm, cons, nss = _split_docs(docs)
        orc, drv, _ = make_pair(tm, cons, nss, lib_path=lib)
        xs = X.System()
        for d in docs:
            if d.get("kind") == "ExpansionTemplate":
                xs.upsert(d)
                drv.AddExpansionTemplate(d)
        objs = [d for d in docs if d.get("kind") in ("Deployment", "ReplicaSet", "CronJob", "Pod", "ConfigMap")]
        revs = [D.Review(object=d, source="Original") for d in objs]
        # admission shapes: the generator of a DELETE is the OLD object (getReqObject, policy.go:435-440), and the request's namespace is
        # set on the generator before it is expanded (policy.go:608): the resultant lands in that namespace
        web = next(d for d in objs if d["metadata"]["name"] == "Web")
        revs += [D.Review(object=None, old_object=web, operation="DELETE", source="Original"),
                 D.Review(object=web, operation="CREATE", source="Original", namespace_name="elsewhere"),
                 D.Review(object=web, old_object=web, operation="UPDATE", source="Original", namespace_name="")]
        resp = drv.ReviewBatch(revs, ep)
        want = set()
        expand_errs = {}
        for i, r in enumerate(revs):
            try:
                for x in X.review_with_expansion(orc, xs, k8s.Review(obj=r.object, old=r.old_object, operation=r.operation, source="Original",
                                                                     namespace=r.namespace_name), ep):
                    want.add((i, "%s/%s" % x["constraint"], x["msg"], json.dumps(x["details"], sort_keys=True), x["enforcementAction"],
                              tuple(x["scopedEnforcementActions"]), bool(x.get("autoreject"))))
            except X.ExpansionError as e:
                expand_errs[i] = str(e)
        got = {x for x in engine_results(resp) if x[0] not in expand_errs}
        assert_same(want, got)
        assert any(x[1].endswith("/priv-scoped") and x[4] == "dryrun" and x[5] for x in got), "override on a scoped constraint: the scoped list must stay"
        assert {i for i, e in enumerate(resp.object_errors or []) if e} == set(expand_errs), (resp.object_errors, expand_errs)
        for i, e in expand_errs.items():
            assert e in resp.object_errors[i]
        msgs = [w[2] for w in want]
        assert any(m.startswith("[Implied by expand-deployments] ") for m in msgs)
        assert any(m.startswith("[Implied by expand-jobs] ") for m in msgs) and any(w[4] == "dryrun" for w in want)
        assert resp.viol_bits.shape[0] == len(revs)
    return len(want)


# ------------------------------------------------------------------------------------------ referential constraints (f-4)
def case_referential(lib):
    """data.inventory: (1) the bats "unique labels test" (test.bats:295-303): with good/no_dupe_cm.yaml synced, bad/no_dupe_cm_2.yaml
    is denied, the synced objects themselves are not (a review never collides with its own inventory entry); (2) a sweep of
    Services / ConfigMaps against an inventory of a few hundred objects, engine == oracle result for result, through the host
    objects path and the raw-JSON blob path; (3) RemoveData takes the violation away again."""
    v = golden("referential_vectors.json")["uniquelabel"]
    tmpl = k8s.template_from_yaml_obj(v["template"])
    orc, drv, skipped = make_pair([tmpl], [v["constraint"]], lib_path=lib)
    assert not skipped
    T = "admission.k8s.gatekeeper.sh"
    for d in v["synced"]:
        orc.add_data(d)
        drv.AddData(T, orc.data_path(d), d)
    revs = [D.Review(object=v["denied"])] + [D.Review(object=d) for d in v["synced"]]
    resp = drv.ReviewBatch(revs, k8s.WEBHOOK_EP)
    want = oracle_results(orc, revs, k8s.WEBHOOK_EP)
    assert_same(want, engine_results(resp))
    # the denied ConfigMap, and the synced one of the matched namespace (its twin in the other namespace carries the same value)
    assert {w[0] for w in want} == {0, 1} and {w[2] for w in want} == {"label gatekeeper has duplicate value not_duplicated"}, want
    twin = v["synced"][1]
    drv.RemoveData(T, orc.data_path(twin))
    orc.remove_data(orc.data_path(twin))
    resp = drv.ReviewBatch(revs, k8s.WEBHOOK_EP)
    want = oracle_results(orc, revs, k8s.WEBHOOK_EP)
    assert_same(want, engine_results(resp))
    assert {w[0] for w in want} == {0}, want          # a review never collides with its own inventory entry
    # the object admitted and synced: now the first ConfigMap collides with it as well
    orc.add_data(v["denied"])
    drv.AddData(T, orc.data_path(v["denied"]), v["denied"])
    resp = drv.ReviewBatch(revs, k8s.WEBHOOK_EP)
    want = oracle_results(orc, revs, k8s.WEBHOOK_EP)
    assert_same(want, engine_results(resp))
    assert {w[0] for w in want} == {0, 1}, want
    drv.RemoveData(T, orc.data_path(v["denied"]))
    orc.remove_data(orc.data_path(v["denied"]))
    resp = drv.ReviewBatch(revs, k8s.WEBHOOK_EP)
    assert_same(oracle_results(orc, revs, k8s.WEBHOOK_EP), engine_results(resp))
    assert {r.object for r in resp.results} == {0}

    # (2) unique service selectors + unique labels over a synthetic inventory
    svc_t = [d for r in golden("gator_test_table.json") if r["name"] == "referential constraint with violation" for d in r["docs"]]
    tm, cons, _ = _split_docs(svc_t)
    rng = random.Random(0xF4)
    def svc(i, ns, sel):
        return {"apiVersion": "v1", "kind": "Service", "metadata": {"name": "svc-%d" % i, "namespace": ns}, "spec": {"ports": [{"port": 443}], "selector": sel}}
    def cm(i, ns, val):
        return {"apiVersion": "v1", "kind": "ConfigMap", "metadata": {"name": "cm-%d" % i, "namespace": ns, "labels": {"gatekeeper": val, "i": str(i)}}, "data": {}}
    nss = ["default", "gatekeeper-test-playground", "prod"]
    inv = []
    for i in range(150):
        inv.append(svc(i, rng.choice(nss), {"app": "a%d" % rng.randrange(60), **({"tier": rng.choice(["fe", "be"])} if rng.random() < 0.3 else {})}))
    for i in range(150):
        inv.append(cm(i, rng.choice(nss), "v%d" % rng.randrange(90)))
    inv.append({"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": "prod", "labels": {"gatekeeper": "v3"}}})
    orc, drv, skipped = make_pair(tm + [tmpl], cons + [v["constraint"]], lib_path=lib)
    assert not skipped
    for d in inv:
        orc.add_data(d)
        drv.AddData(T, orc.data_path(d), d)
    objs = rng.sample(inv, 80) + [svc(1000 + i, rng.choice(nss), {"app": "a%d" % rng.randrange(60)}) for i in range(40)] + \
           [cm(1000 + i, rng.choice(nss), "v%d" % rng.randrange(90)) for i in range(40)] + [cm(2000, "prod", "unique-value"), svc(2000, "prod", {})]
    revs = [D.Review(object=o) for o in objs]
    for ep in (k8s.AUDIT_EP, k8s.WEBHOOK_EP):
        resp = drv.ReviewBatch(revs, ep)
        want = oracle_results(orc, revs, ep)
        assert_same(want, engine_results(resp))
        assert len(want) > 40
    blob = W.PyBlob([json.dumps(o).encode() for o in objs])
    got = drv.ReviewBlob(blob, k8s.AUDIT_EP, flags=D.F_MATERIALIZE)
    assert_same(oracle_results(orc, revs, k8s.AUDIT_EP), engine_results(got))
    return len(want)


def case_audit_lazy(lib, n=3000, limit=4):
    """f-1: constraints with one result per violating pair (the lowering proves it: one path to a head of parameters and object-level
    values) are COUNTED from the bitmap; only the objects that can still enter a constraint's list -- the `limit` smallest by
    (namespace, name), ties included -- are evaluated for their messages.  The report must be the oracle's byte for byte: totals,
    per-action totals, the K-smallest lists in emission order.  Ties: objects that share namespace and name; a page of several
    kinds falls back to evaluating every pair."""
    from oracle import audit as OA
    tm, cons = W.config2()
    nss = W.synth_namespaces()
    orc, drv, _ = make_pair(tm, cons, nss, lib_path=lib)
    dump = drv.Dump()
    single = {l.split()[2] for l in dump.splitlines() if l.startswith("constraint ") and "[one result per pair]" in l}
    assert len(single) >= 8, single
    blob = W.synth_objects(7000, n)
    objs = [json.loads(blob.get(i)) for i in range(n)]
    # ties on (namespace, name): copies of the smallest objects with other labels (other results, same identity)
    order = sorted(range(n), key=lambda i: (objs[i]["metadata"].get("namespace", ""), objs[i]["metadata"]["name"]))
    for j, i in enumerate(order[:6]):
        twin = json.loads(json.dumps(objs[i]))
        twin["metadata"]["labels"] = {"copy": str(j)}
        objs.append(twin)
    rnd = random.Random(5)
    rnd.shuffle(objs)
    nsmap = {x["metadata"]["name"]: x for x in nss}

    def check(run, want):
        got = run.report()
        assert got["totalViolations"] == {"%s/%s" % k: v for k, v in want["totals"].items()}
        assert got["totalViolationsPerEnforcementAction"] == want["by_action"]
        for key, lst in want["violations"].items():
            g = got["violations"]["%s/%s" % key]
            w = [{k: v for k, v in sv.items() if not (k in ("namespace", "enforcementActions") and not v)} for sv in lst]
            assert g == w, (key, g[:2], w[:2])
        return got

    want = OA.audit(orc, objs, namespaces=nsmap, limit=limit)
    # (a) host-flattened pages
    run = D.AuditRun(drv, violations_limit=limit)
    keep = []
    for lo in range(0, len(objs), 1000):
        rb = drv.upload([D.Review(object=o, source="Original") for o in objs[lo:lo + 1000]])
        keep.append(rb)
        run.add_batch(rb, k8s.AUDIT_EP)
    got = check(run, want)
    assert got["pairsCounted"] > 0.5 * sum(want["totals"][k] for k in want["totals"] if "%s/%s" % k in single), got["pairsCounted"]
    # no single-result constraint ever has two results for one pair (what the counting rests on)
    resp = drv.ReviewBatch([D.Review(object=o, source="Original") for o in objs[:800]], k8s.AUDIT_EP)
    seen = {}
    for r in resp.results:
        seen[(r.object, r.constraint)] = seen.get((r.object, r.constraint), 0) + 1
    assert all(v == 1 for (o, c), v in seen.items() if c in single)
    # (b) raw JSON pages through the device ingest
    run = D.AuditRun(drv, violations_limit=limit)
    for lo in range(0, len(objs), 1500):
        pb = W.PyBlob([json.dumps(o).encode() for o in objs[lo:lo + 1500]])
        rb = drv.upload_blob(pb)
        keep.append((pb, rb))
        run.add_batch(rb, k8s.AUDIT_EP)
    got_b = check(run, want)
    assert got_b["pairsCounted"] > 0      # (other page sizes: other candidates, the same report)
    # most violating pairs have one result; those never reach the host evaluator unless they can enter a list
    assert got["pairsEvaluated"] < 0.5 * (got["pairsEvaluated"] + got["pairsCounted"]), (got["pairsEvaluated"], got["pairsCounted"])
    # (c) a page of several kinds: (group, version, kind) orders results first -- every pair is evaluated
    mixed = objs[:300] + [{"apiVersion": "apps/v1", "kind": "Deployment", "metadata": {"name": "d%d" % i, "namespace": "default"}, "spec": {}} for i in range(5)]
    want_m = OA.audit(orc, mixed, namespaces=nsmap, limit=limit)
    run = D.AuditRun(drv, violations_limit=limit)
    rb = drv.upload([D.Review(object=o, source="Original") for o in mixed])
    run.add_batch(rb, k8s.AUDIT_EP)
    got_m = check(run, want_m)
    assert got_m["pairsCounted"] == 0
    return got


def case_wide_objects(lib, pods=128, containers=400):
    """Objects that iterate very many rows: every bit column of an evaluation tile lives in the CTA's shared memory, so the
    backend shrinks the tile (512 -> 32 objects here) instead of refusing the page."""
    t = golden("templates.json")
    tm = [(t[k]["kind"], t[k]["rego"]) for k in ("allowedrepos", "containerlimits", "psp_privileged")]
    cons = [W._constraint(t["allowedrepos"]["kind"], "repos", match=dict(W.POD), params={"repos": ["gcr.io/", "quay.io/team/"]}),
            W._constraint(t["containerlimits"]["kind"], "limits", match=dict(W.POD), params={"cpu": "2", "memory": "1Gi"}),
            W._constraint(t["psp_privileged"]["kind"], "priv", match=dict(W.POD))]
    orc, drv, skipped = make_pair(tm, cons, lib_path=lib)
    assert not skipped
    rnd = random.Random(11)
    objs = []
    for i in range(pods):
        cs = []
        for j in range(containers if i % 3 else 3):
            c = {"name": "c%d" % j, "image": rnd.choice(["gcr.io/a/b:1", "docker.io/x:latest", "quay.io/team/y:2", "evil.io/z"])}
            if rnd.random() < 0.7:
                c["resources"] = {"limits": {"cpu": rnd.choice(["500m", "1", "4", "250m"]), "memory": rnd.choice(["512Mi", "2Gi", "1Gi"])}}
            if rnd.random() < 0.1:
                c["securityContext"] = {"privileged": rnd.random() < 0.5}
            cs.append(c)
        objs.append({"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "wide-%d" % i, "namespace": "default"}, "spec": {"containers": cs}})
    revs = [D.Review(object=o, source="Original") for o in objs]
    want = oracle_results(orc, revs, k8s.AUDIT_EP)
    assert_same(want, engine_results(drv.ReviewBatch(revs, k8s.AUDIT_EP)))
    blob = W.PyBlob([json.dumps(o).encode() for o in objs])
    assert_same(want, engine_results(drv.ReviewBlob(blob, k8s.AUDIT_EP, flags=D.F_MATERIALIZE)))
    return len(want)


def case_audit_concurrent_with_reviews(lib, rounds=6):
    """The audit switches the backend between the decision netlist and the ambiguity netlist (one program resident at a time)
    while other threads review: every review must see its own snapshot's results, and the audit report must not change."""
    import threading
    tm, cons = W.config2()
    nss = W.synth_namespaces()
    orc, drv, _ = make_pair(tm, cons, nss, lib_path=lib)
    blob = W.synth_objects(9000, 600)
    objs = [json.loads(blob.get(i)) for i in range(600)]
    revs = [D.Review(object=o, source="Original") for o in objs[:200]]
    base = engine_results(drv.ReviewBatch(revs, k8s.AUDIT_EP))
    rb = drv.upload([D.Review(object=o, source="Original") for o in objs])
    run0 = D.AuditRun(drv, violations_limit=5)
    run0.add_batch(rb, k8s.AUDIT_EP)
    want = run0.report()
    errors = []

    def reviewer():
        try:
            for _ in range(rounds * 3):
                assert engine_results(drv.ReviewBatch(revs, k8s.AUDIT_EP)) == base
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))

    def auditor():
        try:
            for _ in range(rounds):
                run = D.AuditRun(drv, violations_limit=5)
                run.add_batch(rb, k8s.AUDIT_EP)
                got = run.report()
                assert got["totalViolations"] == want["totalViolations"] and got["violations"] == want["violations"]
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=reviewer) for _ in range(3)] + [threading.Thread(target=auditor) for _ in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:3]
    assert want["pairsCounted"] > 0
    return want


def _spec_env(**kv):
    for k in ("GK_SPEC", "GK_SPEC_MIN_OBJECTS", "GK_SPEC_CHECK"):
        os.environ.pop(k, None)
    os.environ.update(kv)


def case_spec_kernel(lib, n=6000, config=2, wide_every=97):
    """The kernel GENERATED for a constraint set (spec_codegen.cpp -> NVRTC: one thread per object, a mask register per netlist
    node) against the netlist interpreter on the same resident page, through the decision netlist and the audit's ambiguity
    netlist.  Every `wide_every`-th object gets 40 containers: more rows than a mask register holds, so its tile goes to the
    interpreter -- the hand-over is part of the comparison.  On the GPU: two engines (GK_SPEC=0 / forced) must give identical
    bitmaps, error planes, totals and audit reports.  On the TEST-ONLY host emulation: the generated text is compiled with g++
    and checked object by object against the interpreted netlist (GK_SPEC_CHECK=1, tests/_hostemu/hostemu.cpp)."""
    from oracle import audit as OA  # noqa: F401
    if config == 2:
        tm, cons = W.config2()
        mode = 0
    elif config == 4:
        tm, cons = W.config4()
        mode = 1
    else:
        tm, cons = W.config5()
        mode = 0
    nss = W.synth_namespaces()
    blob = W.synth_objects(31337, n, mode=mode)
    docs = [blob.get(i) for i in range(n)]
    nwide = 0
    for i in range(0, n, wide_every):
        d = json.loads(docs[i])
        spec = d.get("spec")
        if not isinstance(spec, dict) or not spec.get("containers"):
            continue
        cs = spec["containers"]
        spec["containers"] = [dict(cs[k % len(cs)], name="c%d" % k) for k in range(40)]
        docs[i] = json.dumps(d).encode()
        nwide += 1
    assert nwide > 3
    page = W.PyBlob(docs)
    is_emu = lib is not None and "hostemu" in lib

    def run(env):
        _spec_env(**env)
        try:
            drv = D.Driver(lib_path=lib)
            for k, r in tm:
                drv.add_template(k, r)
            for c in cons:
                drv.AddConstraint(c)
            for ns in nss:
                drv.AddData("admission.k8s.gatekeeper.sh", ["cluster", "v1", "Namespace", ns["metadata"]["name"]], ns)
            rb = drv.upload_blob(page)
            r = rb.eval(k8s.AUDIT_EP)
            kern = drv.last_kernel()
            w = rb.eval(k8s.WEBHOOK_EP)     # another enforcement-point mask over the same compiled kernel
            au = D.AuditRun(drv, violations_limit=5)
            au.add_batch(rb)                # (evaluates the ambiguity netlist on a fork of the batch: ACC2 through the generated kernel)
            rep = au.report()
            au.close()
            out = (np.array(r.viol_bits, copy=True), np.array(r.err_bits, copy=True), list(r.totals), list(r.err_totals),
                   np.array(w.viol_bits, copy=True), list(w.totals), json.dumps(rep, sort_keys=True))
            rb.free()
            drv.close()
            return kern, out
        finally:
            _spec_env()

    if is_emu:
        kern, out = run({"GK_SPEC_CHECK": "1"})
        return int(sum(out[2]))
    k0, want = run({"GK_SPEC": "0"})
    k1, got = run({"GK_SPEC_MIN_OBJECTS": "0"})
    assert k0 == "gk_eval_kernel" and k1 == "gk_spec_kernel", (k0, k1)
    assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1]), "bitmaps of the generated kernel differ from the interpreter's"
    assert want[2] == got[2] and want[3] == got[3] and np.array_equal(want[4], got[4]) and want[5] == got[5]
    assert want[6] == got[6], "audit reports differ"
    return int(sum(want[2]))
