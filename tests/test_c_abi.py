"""tests/c_abi/mirror.c -- the Go shim's reviewBatch repeated call for call in C (this image has no Go toolchain) -- against the
admission-shape vectors: UPDATE / DELETE with old objects, explicit / cached / missing Namespace, AdmissionRequest.Namespace,
operation, userInfo, every source value.  CPU: linked against the test backend's library (same C ABI); `-m gpu`: the product."""
import json
import os
import subprocess

import pytest

from conftest import HOSTEMU, ROOT, golden, has_cuda
from gatekeeper_b200 import driver as D
from oracle import k8s

SRC = os.path.join(ROOT, "tests", "c_abi", "mirror.c")
BIN = os.path.join(ROOT, "build", "c_abi_mirror")


def _build():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O1", "-Wall", "-o", BIN, SRC, "-ldl"], check=True)


def _blob(x):
    return b"" if x is None else (x if isinstance(x, bytes) else json.dumps(x, separators=(",", ":")).encode())


def _run(lib, tmp_path):
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
    pod = lambda ns, name="p", gen=None: {"apiVersion": "v1", "kind": "Pod",
                                          "metadata": dict({"name": name, "namespace": ns}, **({"generateName": gen} if gen else {}))}
    revs = [
        D.Review(object=pod("a"), old_object=pod("b"), operation="UPDATE", source="Original"),
        D.Review(object=pod("b"), old_object=pod("a"), operation="UPDATE", source="Original"),
        D.Review(object=None, old_object=pod("a"), operation="DELETE", source="Original"),
        D.Review(object=pod("cached"), source="Generated", operation="CREATE"),
        D.Review(object=pod("uncached"), source="Original"),
        D.Review(object=pod("x"), namespace={"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": "x", "labels": {"bar": "qux"}}}, source="Original"),
        D.Review(object=pod("b", name="web-1"), source="Original", user_info={"username": "alice"}),
        D.Review(object=pod("b", name="", gen="web-"), source="Original", user_info={"username": "system:serviceaccount:x"}),
        D.Review(object={"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": "nsobj", "labels": {"bar": "qux"}}}, source="Original"),
        D.Review(object=pod(""), namespace_name="cached", source="Original"),      # AdmissionRequest.Namespace decides the cache lookup
        D.Review(object=pod("b"), source=""),                                      # source unset vs matcher Generated => error
        D.Review(object=b"{not json", source="Original"),                          # review-level error: reported, not dropped
    ]
    src_code = {"": 0, "Original": 1, "Generated": 2, "All": 3}
    path = os.path.join(str(tmp_path), "in.txt")
    with open(path, "wb") as f:
        for kind, rego in tm:
            b = rego.encode()
            f.write(b"T %s %d\n" % (kind.encode(), len(b)) + b + b"\n")
        for c in cons:
            b = json.dumps(c).encode()
            f.write(b"C %d\n" % len(b) + b + b"\n")
        b = json.dumps(ns_qux).encode()
        f.write(b"N cached %d\n" % len(b) + b + b"\n")
        f.write(b"E %s\n" % k8s.WEBHOOK_EP.encode())
        for r in revs:
            o, ol, ns, ui = _blob(r.object), _blob(r.old_object), _blob(r.namespace), _blob(r.user_info or None)
            f.write(b"R %d %s %s %d %d %d %d\n" % (src_code.get(r.source, 4), (r.operation or "-").encode(), (r.namespace_name or "-").encode(),
                                                   len(o), len(ol), len(ns), len(ui)) + o + ol + ns + ui + b"\n")
    out = subprocess.run([BIN, lib, path], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got, errs = set(), {}
    for line in out.stdout.splitlines():
        f = line.split("\t")
        if f[0] == "ERR":
            errs[int(f[1])] = f[2]
        else:
            got.add((int(f[0]), f[1], f[2], f[3] == "1", f[4]))
    # the oracle on the same reviews
    orc = k8s.Client()
    for kind, rego in tm:
        orc.add_template(kind, rego)
    for c in cons:
        orc.add_constraint(c)
    orc.add_namespace(ns_qux)
    want = set()
    for i, r in enumerate(revs[:-1]):
        rv = k8s.Review(obj=r.object, old=r.old_object, ns=r.namespace, source=r.source, operation=r.operation, user_info=r.user_info,
                        namespace=r.namespace_name)
        for x in orc.review(rv, k8s.WEBHOOK_EP):
            want.add((i, "%s/%s" % x["constraint"], x["enforcementAction"], bool(x.get("autoreject")), x["msg"]))
    assert got == want, (sorted(got - want)[:5], sorted(want - got)[:5])
    assert list(errs) == [len(revs) - 1] and "invalid request object" in errs[len(revs) - 1]
    return len(got)


def _run_expansion(lib, tmp_path):
    """UpsertExpansionTemplate / ExpansionConflicts of the shim: a Deployment is reviewed through its resultant Pod ("[Implied by ...]"),
    a pair of templates that expand into each other is stored, reported ("template forms expansion cycle") and listed as conflicts."""
    t = golden("templates.json")
    tmpl = lambda name, group, kind, gen: {"apiVersion": "expansion.gatekeeper.sh/v1beta1", "kind": "ExpansionTemplate", "metadata": {"name": name},
                                            "spec": {"applyTo": [{"groups": [group], "versions": ["v1"], "kinds": [kind]}], "templateSource": "spec.template",
                                                     "generatedGVK": {"group": gen[0], "version": "v1", "kind": gen[1]}}}
    xts = [tmpl("expand-deployments", "apps", "Deployment", ("", "Pod")), tmpl("a-to-b", "ga", "A", ("gb", "B")), tmpl("b-to-a", "gb", "B", ("ga", "A"))]
    con = {"kind": "NeverValidate", "metadata": {"name": "pods"}, "spec": {"match": {"kinds": [{"apiGroups": [""], "kinds": ["Pod"]}]}}}
    dep = {"apiVersion": "apps/v1", "kind": "Deployment", "metadata": {"name": "web", "namespace": "a"}, "spec": {"template": {"metadata": {"labels": {"x": "y"}}}}}
    path = os.path.join(str(tmp_path), "in_x.txt")
    with open(path, "wb") as f:
        b = t["fixtures_TemplateNeverValidate"]["rego"].encode()
        f.write(b"T NeverValidate %d\n" % len(b) + b + b"\n")
        b = json.dumps(con).encode()
        f.write(b"C %d\n" % len(b) + b + b"\n")
        for x in xts:
            b = json.dumps(x).encode()
            f.write(b"X %d\n" % len(b) + b + b"\n")
        f.write(b"E %s\n" % k8s.AUDIT_EP.encode())
        o = _blob(dep)
        f.write(b"R 1 - - %d 0 0 0\n" % len(o) + o + b"\n")
    out = subprocess.run([BIN, lib, path], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = [ln.split("\t") for ln in out.stdout.splitlines()]
    assert [ln[1] for ln in lines if ln[0] == "XERR"] == ["template forms expansion cycle"], lines
    assert [json.loads(ln[1]) for ln in lines if ln[0] == "CONFLICTS"] == [["a-to-b", "b-to-a"]], lines
    res = [ln for ln in lines if ln[0] == "0"]
    assert len(res) == 1 and res[0][1] == "NeverValidate/pods" and res[0][4].startswith("[Implied by expand-deployments] "), lines
    return len(res)


def test_c_mirror_of_the_go_shim_on_the_test_backend(tmp_path):
    _build()
    assert _run(HOSTEMU, tmp_path) >= 10
    assert _run_expansion(HOSTEMU, tmp_path) == 1


@pytest.mark.gpu
def test_c_mirror_of_the_go_shim_on_the_cuda_library(tmp_path):
    if not has_cuda():
        pytest.fail("GPU tests selected but no CUDA device is visible")
    _build()
    assert _run(os.path.join(ROOT, "gatekeeper_b200", "libgk_engine.so"), tmp_path) >= 10
