import os, sys, time
sys.path.insert(0, os.getcwd())
from gatekeeper_b200 import driver as D, workloads as W
n=int(os.environ.get("N", "400000"))
blob = W.synth_objects(0, n)
tm, cons = W.config2()
for th in (1, 16):
    drv = D.Driver(threads=th)
    for k, r in tm: drv.add_template(k, r)
    for c in cons: drv.AddConstraint(c)
    for ns in W.synth_namespaces(): drv.AddData("t", ["cluster", "v1", "Namespace", ns["metadata"]["name"]], ns)
    best=1e9
    for rep in range(2):
        t0 = time.perf_counter(); rb = drv.upload_blob(blob); dt = time.perf_counter() - t0
        best=min(best, dt); fl=rb.stats["flatten_ms"]; rb.free()
    print("threads %3d: upload %.2f s (flatten %.0f ms)  %.2f us/obj  cpu-us/obj %.1f" % (th, best, fl, 1e6*best/n, 1e6*best/n*th), flush=True)
    drv.close()
