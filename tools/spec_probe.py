"""GPU probe: the kernel generated for the constraint set (gk_spec_kernel, NVRTC) against the netlist interpreter
(gk_eval_kernel) on a resident page of a BASELINE configuration -- identical bitmaps / totals, and the launch time of
several build variants (threads per CTA x minimum resident CTAs = register cap).

    python tools/spec_probe.py [--config 2] [--objects 1000000] [--variants 512x1,512x1:GENERIC,256x2,384x1,512x1:NOMATCH]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--objects", type=int, default=1000000)
ap.add_argument("--variants", default="512x1,512x1:GENERIC,256x2,384x1,512x1:NOMATCH")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--wide", type=int, default=0, help="every WIDE-th object gets 40 containers (exercises the tile hand-over)")
a = ap.parse_args()

from gatekeeper_b200 import driver as D, workloads as Wl  # noqa: E402


def engine():
    drv = D.Driver()
    if a.config == 2:
        tmpls, cons = Wl.config2()
    elif a.config == 4:
        tmpls, cons = Wl.config4()
    else:
        tmpls, cons = Wl.config5()
    for kind, rego in tmpls:
        drv.add_template(kind, rego)
    for c in cons:
        drv.AddConstraint(c)
    for ns in Wl.synth_namespaces():
        drv.AddData("admission.k8s.gatekeeper.sh", ["cluster", "v1", "Namespace", ns["metadata"]["name"]], ns)
    return drv


mode = 1 if a.config == 4 else 0
blob = Wl.synth_objects(0, a.objects, mode=mode)
if a.wide:
    docs = [blob.get(i) for i in range(a.objects)]
    for i in range(0, a.objects, a.wide):
        d = json.loads(docs[i])
        cs = d.get("spec", {}).get("containers") or [{"name": "c", "image": "nginx"}]
        d["spec"]["containers"] = [dict(cs[k % len(cs)], name="c%d" % k) for k in range(40)]
        docs[i] = json.dumps(d).encode()
    blob = Wl.PyBlob(docs)


def run(label, env):
    for k in ("GK_SPEC", "GK_SPEC_THREADS", "GK_SPEC_MINB", "GK_SPEC_DEFS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    drv = engine()
    drv.pin_blob(blob)
    rb = drv.upload_blob(blob)
    t0 = time.perf_counter()
    r = rb.eval(D.AUDIT_EP)            # (the first evaluation builds the kernel: outside the timed launches below)
    first_s = time.perf_counter() - t0
    ms = []
    for _ in range(3 + a.steps):
        ms.append(rb.eval(D.AUDIT_EP, D.F_NO_COPY_BACK).stats["kernel_ms"])
    ms = ms[3:]
    out = {"label": label, "kernel": drv.last_kernel(), "kernel_ms": round(sum(ms) / len(ms), 4), "min_ms": round(min(ms), 4), "first_eval_s": round(first_s, 2),
           "alg_bytes": rb.alg_bytes + a.objects * r.viol_bits.shape[1] * 8}
    out["GBps"] = round(out["alg_bytes"] / out["kernel_ms"] / 1e6, 1)
    res = (np.array(r.viol_bits, copy=True), np.array(r.err_bits, copy=True), list(r.totals), list(r.err_totals))
    rb.free()
    drv.pin_blob(blob, False)
    drv.close()
    return out, res


base, want = run("interpreter", {"GK_SPEC": "0"})
print(json.dumps(base), flush=True)
for v in a.variants.split(","):     # THREADSxMINB[:SWITCH+SWITCH]: GENERIC = spec.match through the shared gk_match(); NOMATCH = -DGK_SPEC_X_NOMATCH (results differ)
    v, _, sw = v.partition(":")
    th, mb = v.split("x")
    toks = [x for x in sw.split("+") if x]
    defs = " ".join("-DGK_SPEC_X_" + x for x in toks if x != "GENERIC")
    os.environ["GK_SPEC_MATCHGEN"] = "0" if "GENERIC" in toks else "1"
    o, got = run("spec " + v + (" " + sw if sw else ""), {"GK_SPEC_THREADS": th, "GK_SPEC_MINB": mb, "GK_SPEC_MIN_OBJECTS": "0", "GK_SPEC_DEFS": defs})
    o["identical"] = bool(np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1]) and want[2] == got[2] and want[3] == got[3])
    print(json.dumps(o), flush=True)
    assert o["identical"] or "NOMATCH" in sw, "generated kernel and interpreter disagree"
    assert o["kernel"] == "gk_spec_kernel" or "hostemu" in os.environ.get("GK_ENGINE_LIB", ""), "the generated kernel did not run"
