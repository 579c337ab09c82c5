#!/usr/bin/env python
"""Admission Review replay (BASELINE.json configs[2]): 200 constraints = the 5 PSP constraints cloned round-robin
(generateConstraints, pkg/webhook/policy_benchmark_test.go:191-199), UPDATE requests carrying object + oldObject like
createAdmissionRequests (:210-249), reviewed in micro-batches of 64 through the public API (host JSON in, deny/warn
messages out).  Latency of a request = latency of its micro-batch; percentiles as `gator bench` defines them
(pkg/gator/bench/metrics.go:9-59).

  python tools/admission_latency.py [batches] [batch_size]
(the CPU restatement of the same replay is timed by tests/admission_cpu_baseline.py -- the oracle is test infrastructure)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gatekeeper_b200 import driver as D
from gatekeeper_b200 import metrics as M
from gatekeeper_b200 import workloads as W


def requests(pods, n):
    return [D.Review(object=pods[i % len(pods)], old_object=pods[(i + 1) % len(pods)], operation="UPDATE",
                     namespace_name=pods[i % len(pods)]["metadata"].get("namespace"), user_info={"username": "bench"}) for i in range(n)]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    batches = int(args[0]) if args else 200
    bs = int(args[1]) if len(args) > 1 else 64
    tm, cons, pods = W.config3(200)
    drv = D.Driver()
    for k, r in tm:
        drv.add_template(k, r)
    for x in cons:
        drv.AddConstraint(x)
    revs = requests(pods, bs)
    # the timed region is the C ABI: gk_review_batch (flatten + H2D + kernel + D2H + message rendering) and one
    # gk_validation_messages per request; request marshalling (Python dict -> JSON bytes) happens once, outside
    import ctypes as C
    arr, n, keep = drv._marshal(revs)
    lib, eng = drv._lib, drv._e
    flags = D.F_MATERIALIZE | D.F_PROCESS_WEBHOOK

    def one_batch(count_denied=False):
        res = D.gk_result()
        err = C.c_char_p()
        drv._check(lib.gk_review_batch(eng, arr, n, D.WEBHOOK_EP.encode(), flags, C.byref(res), C.byref(err)), err)
        denied = 0
        for i in range(n):
            p = lib.gk_validation_messages(eng, C.byref(res), i, C.byref(err))
            if count_denied and b'"deny":[]' not in C.string_at(p):
                denied += 1
            lib.gk_free_str(p)
        stats = {k: getattr(res, k) for k in ("flatten_ms", "h2d_ms", "kernel_ms", "d2h_ms", "materialize_ms")}
        stats["results"] = res.n_violations
        lib.gk_free_result(C.byref(res))
        return denied, stats

    for _ in range(5):
        one_batch()
    _, breakdown = one_batch()
    breakdown = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in breakdown.items()}
    lat = []
    t_all = time.perf_counter_ns()
    denied = 0
    for _ in range(batches):
        t0 = time.perf_counter_ns()
        d, _s = one_batch(count_denied=True)
        dt = time.perf_counter_ns() - t0
        lat.extend([dt] * bs)
        denied += d
    total = time.perf_counter_ns() - t_all
    out = {"impl": "gatekeeper_b200 (" + drv.backend() + ")", "micro_batch": bs, "batches": batches, "constraints": len(cons),
           "latency_ns": M.calculate_latencies(lat), "requests_per_s": M.calculate_throughput(batches * bs, total),
           "evals_per_s": M.calculate_throughput(batches * bs, total) * len(cons), "denied_requests": denied,
           "engine_breakdown_of_one_batch_ms": breakdown}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
