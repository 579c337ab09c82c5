#!/usr/bin/env python
"""Kernel-only timing of the resident config-2 batch (CUDA events inside the engine), for kernel experiments.
  GK_ENGINE_LIB=gatekeeper_b200/libgk_engine_tXXX.so python tools/kernel_probe.py [objects] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gatekeeper_b200 import driver as D
from gatekeeper_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
tm, cons = W.config2()
drv = D.Driver()
for k, r in tm:
    drv.add_template(k, r)
for c in cons:
    drv.AddConstraint(c)
for ns in W.synth_namespaces():
    drv.AddData("t", ["cluster", "v1", "Namespace", ns["metadata"]["name"]], ns)
blob = W.synth_objects(0, n)
t0 = time.time()
rb = drv.upload_blob(blob)
up = time.time() - t0
ms = []
for _ in range(reps + 3):
    ms.append(rb.eval(flags=D.F_NO_COPY_BACK).stats["kernel_ms"])
ms = sorted(ms[3:])
alg = rb.alg_bytes + n * 2 * 4 * 2
print(f"{os.environ.get('GK_ENGINE_LIB', 'default')} [{drv.last_kernel()}]: n={n} upload {up:.1f}s flatten {rb.stats['flatten_ms']:.0f}ms kernel median {ms[len(ms)//2]:.3f} ms "
      f"min {ms[0]:.3f} ms  -> {n * 50 / ms[len(ms)//2] / 1e6:.1f} G evals/s, {alg / ms[len(ms)//2] / 1e6:.0f} GB/s")
