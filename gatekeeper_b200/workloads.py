"""Synthetic workloads of BASELINE.json's configs: constraint sets built from the
reference's in-tree templates (tests/golden/templates.json) + deterministic objects from libgk_synth.so."""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SYNTH_PATH = os.path.join(_HERE, "libgk_synth.so")
SEED = 0x6A7E6B33

_templates = None


def templates() -> dict:
    global _templates
    if _templates is None:
        with open(os.path.join(ROOT, "tests", "golden", "templates.json")) as f:
            _templates = json.load(f)
    return _templates


# ------------------------------------------------------------------------------------------ synthetic objects
_lib = None


def _synth():
    global _lib
    if _lib is None:
        if not os.path.exists(SYNTH_PATH):
            raise RuntimeError(f"{SYNTH_PATH} missing: run __graft_entry__.build()")
        _lib = C.CDLL(SYNTH_PATH)
        _lib.gk_synth_objects.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_void_p),
                                          C.POINTER(C.POINTER(C.c_uint64))]
        _lib.gk_synth_namespaces.argtypes = [C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint64)]
        _lib.gk_synth_free.argtypes = [C.c_void_p]
    return _lib


class ObjectBlob:
    """`count` JSON documents in one contiguous host buffer (kept alive while reviews point into it)."""

    def __init__(self, buf, offsets, count):
        self.buf, self.offsets, self.count = buf, offsets, count

    def __len__(self):
        return self.count

    def get(self, i) -> bytes:
        return C.string_at(self.buf.value + self.offsets[i], self.offsets[i + 1] - self.offsets[i])

    def total_bytes(self) -> int:
        return int(self.offsets[self.count])

    def __del__(self):
        try:
            lib = _synth()
            lib.gk_synth_free(self.buf)
            lib.gk_synth_free(C.cast(self.offsets, C.c_void_p))
        except Exception:
            pass


def synth_objects(start: int, count: int, mode: int = 0, seed: int = SEED, threads: int = 0) -> ObjectBlob:
    lib = _synth()
    buf = C.c_void_p()
    off = C.POINTER(C.c_uint64)()
    rc = lib.gk_synth_objects(seed, start, count, mode, threads, C.byref(buf), C.byref(off))
    if rc != 0:
        raise RuntimeError(f"gk_synth_objects failed: {rc}")
    return ObjectBlob(buf, off, count)


def synth_namespaces(seed: int = SEED) -> List[dict]:
    lib = _synth()
    buf = C.c_void_p()
    off = C.POINTER(C.c_uint64)()
    n = C.c_uint64()
    rc = lib.gk_synth_namespaces(seed, C.byref(buf), C.byref(off), C.byref(n))
    if rc != 0:
        raise RuntimeError("gk_synth_namespaces failed")
    out = [json.loads(C.string_at(buf.value + off[i], off[i + 1] - off[i])) for i in range(n.value)]
    lib.gk_synth_free(buf)
    lib.gk_synth_free(C.cast(off, C.c_void_p))
    return out


# ------------------------------------------------------------------------------------------ constraint sets
def _constraint(kind, name, match=None, params=None, action=None, scoped=None):
    spec = {}
    if match is not None:
        spec["match"] = match
    if params is not None:
        spec["parameters"] = params
    if action:
        spec["enforcementAction"] = action
    if scoped:
        spec["scopedEnforcementActions"] = scoped
    return {"apiVersion": "constraints.gatekeeper.sh/v1beta1", "kind": kind, "metadata": {"name": name}, "spec": spec}


POD = {"kinds": [{"apiGroups": [""], "kinds": ["Pod"]}]}


def match_variants():
    """The spec.match shapes SURVEY.md 8(d) asks config 2 to mix."""
    return [
        dict(POD),
        dict(POD, namespaces=["prod-*", "ns-00*", "production"]),
        dict(POD, excludedNamespaces=["kube-*", "*-system"]),
        dict(POD, labelSelector={"matchLabels": {"env": "prod"},
                                 "matchExpressions": [{"key": "tier", "operator": "NotIn", "values": ["v1", "v2"]},
                                                      {"key": "team", "operator": "Exists"}]}),
        dict(POD, scope="Namespaced", namespaceSelector={"matchExpressions": [{"key": "env", "operator": "In", "values": ["prod", "staging"]}]}),
    ]


CONFIG2_TEMPLATES = ["requiredlabels_basic", "allowedrepos_prefixmatch", "containerlimits", "requiredprobes", "bannedimagetags",
                     "psp_privileged", "psp_hostnamespace", "psp_hostnetworkports", "psp_volumetypes", "psp_hostfilesystem"]

CONFIG2_PARAMS = {
    "requiredlabels_basic": [{"labels": ["team"]}, {"labels": ["team", "app"]}, {"labels": ["owner"]}, {"labels": ["env", "tier"]},
                             {"labels": ["label-07"]}],
    "allowedrepos_prefixmatch": [{"repos": ["openpolicyagent/"]}, {"repos": ["gcr.io/", "docker.io/library/"]},
                                 {"repos": ["gcr.io/proj-0", "quay.io/"]},
                                 {"repos": ["openpolicyagent/", "gcr.io/", "docker.io/library/", "quay.io/", "registry.k8s.io/"]},
                                 {"repos": ["registry.k8s.io/"]}],
    "containerlimits": [{"cpu": "500m", "memory": "512Mi"}, {"cpu": "1", "memory": "1Gi"}, {"cpu": "2", "memory": "2Gi"},
                        {"cpu": "250m", "memory": "128Mi"}, {"cpu": "4", "memory": "4Gi"}],
    "requiredprobes": [{"probes": ["readinessProbe", "livenessProbe"], "probeTypes": ["tcpSocket", "httpGet", "exec"]},
                       {"probes": ["readinessProbe"], "probeTypes": ["httpGet"]},
                       {"probes": ["livenessProbe"], "probeTypes": ["tcpSocket", "exec"]},
                       {"probes": ["readinessProbe", "livenessProbe", "startupProbe"], "probeTypes": ["tcpSocket", "httpGet", "exec"]},
                       {"probes": ["livenessProbe"], "probeTypes": ["httpGet"]}],
    "bannedimagetags": [{"tags": ["latest"]}, {"tags": ["latest", "v0.0.1"]}, {"tags": ["v1.2.3", "v3.9.19"]}, {"tags": ["latest", "v0.0.0", "v0.0.1", "v0.0.2"]},
                        {"tags": ["sha-00000000"]}],
    "psp_privileged": [None, None, None, None, None],
    "psp_hostnamespace": [None, None, None, None, None],
    "psp_hostnetworkports": [{"hostNetwork": True, "min": 80, "max": 9000}, {"hostNetwork": False, "min": 1024, "max": 65535},
                             {"hostNetwork": False, "min": 0, "max": 0}, {"hostNetwork": True, "min": 30000, "max": 32767},
                             {"min": 8000, "max": 8999}],
    "psp_volumetypes": [{"volumes": ["configMap", "emptyDir", "projected", "secret", "downwardAPI", "persistentVolumeClaim", "flexVolume"]},
                        {"volumes": ["*"]}, {"volumes": ["configMap", "secret"]}, {"volumes": ["emptyDir", "hostPath", "nfs"]},
                        {"volumes": []}],
    "psp_hostfilesystem": [{"allowedHostPaths": [{"pathPrefix": "/foo", "readOnly": True}]}, {"allowedHostPaths": []},
                           {"allowedHostPaths": [{"pathPrefix": "/var/log"}, {"pathPrefix": "/tmp", "readOnly": True}]},
                           {"allowedHostPaths": [{"pathPrefix": "/", "readOnly": True}]},
                           {"allowedHostPaths": [{"pathPrefix": "/etc"}]}],
}


def config2() -> Tuple[List[Tuple[str, str]], List[dict]]:
    """audit sweep: 10 templates x 5 constraints = 50 constraints over Pods."""
    t = templates()
    tmpls = [(t[n]["kind"], t[n]["rego"]) for n in CONFIG2_TEMPLATES]
    cons = []
    mv = match_variants()
    actions = [None, "dryrun", "warn", None,
               ("scoped", [{"action": "deny", "enforcementPoints": [{"name": "audit.gatekeeper.sh"}]},
                           {"action": "warn", "enforcementPoints": [{"name": "*"}]}])]
    for ti, n in enumerate(CONFIG2_TEMPLATES):
        for vi in range(5):
            a = actions[(ti + vi) % 5]
            act, scoped = (a if isinstance(a, tuple) else (a, None))
            cons.append(_constraint(t[n]["kind"], f"{n.replace('_', '-')}-{vi}", match=mv[(ti + vi) % 5], params=CONFIG2_PARAMS[n][vi],
                                    action=act, scoped=scoped))
    return tmpls, cons


def config1():
    """gator verify plumbing: K8sRequiredLabels (basic), 1 constraint labels:[team] on Pods, 100 Pods."""
    t = templates()["requiredlabels_basic"]
    return [(t["kind"], t["rego"])], [_constraint(t["kind"], "must-have-team", match=dict(POD), params={"labels": ["team"]})]


def config3(n_constraints: int = 200, seed: int = 7):
    """admission replay: the 5 PSP constraints cloned round-robin with random 10-letter names
    (generateConstraints, pkg/webhook/policy_benchmark_test.go:191-199)."""
    import random
    with open(os.path.join(ROOT, "tests", "golden", "psp_suite.json")) as f:
        psp = json.load(f)
    rnd = random.Random(seed)
    tmpls = [(x["kind"], x["rego"]) for x in psp["templates"]]
    cons = []
    for i in range(n_constraints):
        base = json.loads(json.dumps(psp["constraints"][i % len(psp["constraints"])]))
        base["metadata"]["name"] = "".join(rnd.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(10))
        cons.append(base)
    return tmpls, cons, psp["pods"]


def config4():
    """K8sPSP* suite x mixed-GVK objects (synth mode 1)."""
    t = templates()
    names = ["psp_privileged", "psp_hostnamespace", "psp_hostnetworkports", "psp_volumetypes", "psp_hostfilesystem"]
    tmpls = [(t[n]["kind"], t[n]["rego"]) for n in names]
    cons = [_constraint(t[n]["kind"], n.replace("_", "-"), match=dict(POD), params=CONFIG2_PARAMS[n][0]) for n in names]
    return tmpls, cons


def config5():
    """K8sAllowedRepos glob/wildcard stress: prefix lists of 1-16 repos + namespaces/name wildcards."""
    t = templates()["allowedrepos_prefixmatch"]
    regs = ["openpolicyagent/", "gcr.io/proj-00/", "gcr.io/proj-01/", "gcr.io/proj-02/", "gcr.io/proj-03/", "docker.io/library/", "quay.io/",
            "registry.k8s.io/", "gcr.io/proj-04/", "gcr.io/proj-05/", "gcr.io/proj-06/", "gcr.io/proj-07/", "gcr.io/proj-08/", "gcr.io/proj-09/",
            "gcr.io/proj-10/", "gcr.io/proj-11/"]
    cons = []
    pats = [None, {"namespaces": ["ns-0*"]}, {"excludedNamespaces": ["*-system", "ns-09*"]}, {"name": "pod-1*"}, {"name": "*-00*"},
            {"namespaces": ["*9*"], "name": "*a"}, {"namespaces": ["production", "ns-0001", "ns-0002*"]}, {"excludedNamespaces": ["*"]}]
    for i, k in enumerate([1, 2, 3, 4, 6, 8, 12, 16]):
        m = dict(POD)
        if pats[i]:
            m.update(pats[i])
        cons.append(_constraint(t["kind"], f"repos-{k}", match=m, params={"repos": regs[:k]}))
    return [(t["kind"], t["rego"])], cons


class PyBlob(ObjectBlob):
    """An ObjectBlob over Python-owned memory: the JSON documents (bytes, or objects to be serialised) back to back."""

    def __init__(self, docs):
        raw = [d if isinstance(d, (bytes, bytearray)) else json.dumps(d, separators=(",", ":")).encode() for d in docs]
        self._data = b"".join(raw)
        self._cbuf = C.create_string_buffer(self._data, len(self._data) + 1)
        off = (C.c_uint64 * (len(raw) + 1))()
        pos = 0
        for i, r in enumerate(raw):
            off[i] = pos
            pos += len(r)
        off[len(raw)] = pos
        self.buf = C.c_void_p(C.addressof(self._cbuf))
        self.offsets = off
        self.count = len(raw)

    def __del__(self):
        pass
