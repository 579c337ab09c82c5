#!/usr/bin/env python
"""Stage timing of the device ingest path on the config-2 workload:  python tools/ingest_probe.py [objects] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gatekeeper_b200 import driver as D
from gatekeeper_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tm, cons = W.config2()
drv = D.Driver()
for k, r in tm:
    drv.add_template(k, r)
for c in cons:
    drv.AddConstraint(c)
for ns in W.synth_namespaces():
    drv.AddData("t", ["cluster", "v1", "Namespace", ns["metadata"]["name"]], ns)
print(drv.Dump().splitlines()[1])
blob = W.synth_objects(0, n)
print("blob bytes", blob.total_bytes())
if os.environ.get("GK_PIN", "1") == "1":
    drv.pin_blob(blob)
blob2 = W.synth_objects(n, n)
if os.environ.get("GK_PIN", "1") == "1":
    drv.pin_blob(blob2)
pages = [blob, blob2]
pipelined = os.environ.get("GK_PIPELINE", "1") == "1"
if pipelined:
    drv.prefetch_blob(pages[0])
for r in range(reps):
    t0 = time.time()
    if pipelined and r + 1 < reps:
        drv.prefetch_blob(pages[(r + 1) % 2])
    resp = drv.ReviewBlob(pages[r % 2], flags=D.F_NO_COPY_BACK, with_results=False)
    dt = time.time() - t0
    print(f"rep {r}: wall {dt*1e3:.1f} ms -> {n*50/dt/1e6:.1f} M evals/s; stats {resp.stats}")
if os.environ.get("GK_TRACE_INGEST"):
    pass
