#!/usr/bin/env python
"""CPU baseline of the admission replay (BASELINE.json configs[2]) on the ORACLE (test infrastructure): 200 constraints,
one request at a time like pkg/webhook/policy.go:661, latency percentiles as gator bench defines them.  Not a test.

  python tests/admission_cpu_baseline.py [requests]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gatekeeper_b200 import metrics as M
from gatekeeper_b200 import workloads as W
from oracle import k8s


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    tm, cons, pods = W.config3(200)
    c = k8s.Client()
    for k, r in tm:
        c.add_template(k, r)
    for x in cons:
        c.add_constraint(x)
    lat = []
    for i in range(n):
        t0 = time.perf_counter_ns()
        c.review(k8s.Review(obj=pods[i % 5], old=pods[(i + 1) % 5], operation="UPDATE"), k8s.WEBHOOK_EP)
        lat.append(time.perf_counter_ns() - t0)
    print(json.dumps({"impl": "oracle (CPU restatement, 1 core)", "requests": n, "constraints": len(cons), "latency_ns": M.calculate_latencies(lat),
                      "requests_per_s": M.calculate_throughput(n, sum(lat))}))


if __name__ == "__main__":
    main()
