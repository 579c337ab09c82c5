#!/usr/bin/env python
"""Benchmark of the audit-sweep hot path (BASELINE.json: constraint x object evaluations / second).

  python bench.py --gpus N --steps K --warmup W              # this framework, one rank per GPU (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle port) on the host cores

Workload (config.workload): BASELINE.json configs[1] -- "audit sweep: 10 templates, 50 constraints, 1M synthetic
Pods on 1 B200".  Weak scaling: every rank evaluates its own 1M-Pod shard of the cluster and the violation
bitmaps are all-gathered once per step (the one exchange the path has).

One step = one pass of the hot path over the batch: the fused match-prefilter + predicate kernel over the
column-wise batch resident in HBM, producing the violation/error bitmaps and per-constraint totals, plus (N > 1)
the NCCL all-gather of the bitmap and all-reduce of the totals.  `value` is device-timed (CUDA events, max over
ranks) with inputs resident; `e2e` goes through the public API with HOST JSON buffers every step (flatten ->
pinned staging -> H2D -> kernel -> D2H of bitmaps + totals).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "constraint_object_evals_per_sec_audit_sweep"
UNIT = "evals/s"
L2_BYTES = 126 * 1024 * 1024


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.lines = []
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ oracle legs
_W = {}


def _oracle_init():
    from gatekeeper_b200 import workloads as W
    from oracle import k8s
    tm, cons = W.config2()
    c = k8s.Client()
    for k, r in tm:
        c.add_template(k, r)
    for x in cons:
        c.add_constraint(x)
    nss = {n["metadata"]["name"]: n for n in W.synth_namespaces()}
    for n in nss.values():
        c.add_namespace(n)
    _W["client"], _W["ns"], _W["k8s"], _W["W"] = c, nss, k8s, W


def _oracle_chunk(rng):
    """Sequential Review of objects [lo, hi) exactly as the reference's audit loop does (pkg/audit/manager.go:686-720):
    JSON decode, namespace lookup, Client.Review.  Returns (#objects, #violations)."""
    if "client" not in _W:
        _oracle_init()
    lo, hi = rng
    W, k8s, c = _W["W"], _W["k8s"], _W["client"]
    blob = W.synth_objects(lo, hi - lo, threads=1)
    nv = 0
    for i in range(hi - lo):
        obj = json.loads(blob.get(i))
        ns = _W["ns"].get(obj["metadata"].get("namespace", ""))
        nv += len(c.review(k8s.Review(obj=obj, ns=ns, source="Original"), k8s.AUDIT_EP))
    return hi - lo, nv


def cpu_baseline_single(sample=2000):
    _oracle_init()
    t0 = time.perf_counter()
    n, nv = _oracle_chunk((0, sample))
    dt = time.perf_counter() - t0
    c = len(_W["client"].constraints)
    return {"value": n * c / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle/ (Python restatement of Client.Review + Rego evaluation), first {n} Pods of the workload x {c} constraints, "
                      f"{dt:.1f} s, sequential like pkg/audit/manager.go:686-720"}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    import multiprocessing as mp
    from gatekeeper_b200.hostinfo import host_cpus
    procs = max(1, host_cpus())                      # every CPU the cgroup lets this process use
    per = args.ref_objects_per_core
    C = 50
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_oracle_init) as pool:
        def step(k):
            base = k * procs * per
            t0 = time.perf_counter()
            res = pool.map(_oracle_chunk, [(base + i * per, base + (i + 1) * per) for i in range(procs)])
            return time.perf_counter() - t0, sum(r[0] for r in res)
        for k in range(args.warmup):
            step(1000 + k)
        total_t, total_n = 0.0, 0
        for k in range(args.steps):
            dt, n = step(k)
            total_t += dt
            total_n += n
    value = total_n * C / total_t
    sample = (f"oracle/ port of the reference CPU path over {procs} processes (= usable CPUs: affinity + cgroup quota of a {os.cpu_count()}-thread host); each step reviews {procs * per} synthetic Pods x {C} constraints "
              f"(a bounded sample of the 1M-Pod workload; objects are independent so throughput is size-independent)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_t / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int",
            "data": "synthetic", "config": {"workload": "audit sweep: 10 templates, 50 constraints, synthetic Pods (configs[1])",
                                            "objects_per_step": procs * per, "constraints": C},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from gatekeeper_b200 import driver as D
    from gatekeeper_b200 import workloads as W

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    tm, cons = W.config2()
    from gatekeeper_b200.hostinfo import host_cpus
    host_threads = max(1, host_cpus() // max(1, world))
    drv = D.Driver(device=local, threads=host_threads)
    for k, r in tm:
        drv.add_template(k, r)
    for c in cons:
        drv.AddConstraint(c)
    for ns in W.synth_namespaces():
        drv.AddData("admission.k8s.gatekeeper.sh", ["cluster", "v1", "Namespace", ns["metadata"]["name"]], ns)
    C = len(drv.constraints())
    words = (C + 31) // 32
    n = args.objects
    t0 = time.perf_counter()
    blob = W.synth_objects(rank * n, n, threads=host_threads)
    gen_s = time.perf_counter() - t0
    rb = drv.upload_blob(blob)                       # flatten + H2D once: inputs resident for the `value` leg
    alg_in = rb.alg_bytes
    alg_bytes = alg_in + n * words * 4 * 2           # + violation and error planes written

    from gatekeeper_b200.sweep import ShardedSweep
    sweep = ShardedSweep(drv, rb, n, C, dev, world)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev) if alg_in < 2 * L2_BYTES else None
    ep = D.AUDIT_EP
    stream = torch.cuda.current_stream()

    def step(e0=None, e1=None):
        if flush is not None:
            flush.fill_(1)                            # evict L2 between iterations when the batch could fit in it
        if e0 is not None:
            e0.record(stream)
        sweep.evaluate(ep, stream)                    # the fused match + predicate kernel over the resident shard
        if e1 is not None:
            e1.record(stream)
        sweep.exchange()                              # N > 1: bitmap all-gather + totals all-reduce (NCCL)

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    launches0 = rb.eval(ep, D.F_NO_COPY_BACK).stats["gpu_launches"]     # engine launch counter before the timed region
    if world > 1:
        dist.barrier()
    K = args.steps
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    ev_a.record(stream)
    for k in range(K):
        step(e0[k], e1[k])
    ev_b.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev_a.elapsed_time(ev_b)
    kern_ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1)) / K
    t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / K
    value = world * n * C / (ms_per_step / 1e3)
    launches = rb.eval(ep, D.F_NO_COPY_BACK).stats["gpu_launches"] - launches0 - 1   # minus this probe's own launch
    totals_host = sweep.tot[0].tolist()              # after the exchange: summed over ranks
    sweep.evaluate(ep, stream)                       # this rank's own totals again (untimed), to check the e2e path against
    torch.cuda.synchronize()
    totals_local = sweep.tot_local[0].tolist()
    if world > 1:
        # the exchange (fused peer stores or NCCL all-gather) must agree with an independent all-reduce of the local totals,
        # and the gathered bitmap with the totals it came with
        ref = sweep.tot_local.clone()
        dist.all_reduce(ref)
        assert torch.equal(ref, sweep.tot), "exchanged totals differ from the all-reduced local totals"
        g = sweep.gathered.contiguous().view(world * n, words)
        bit0 = int(((g[:, 0] & 1) != 0).sum())
        assert bit0 == int(sweep.tot[0, 0]), "gathered bitmap and exchanged totals disagree"

    # ---- e2e: public API, host JSON buffers in, bitmaps + totals out, every step
    e2e_steps = max(1, min(K, args.e2e_steps))
    drv.ReviewBlob(blob, ep)                           # warm-up (pinned staging buffer allocation etc.)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        resp = drv.ReviewBlob(blob, ep)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    assert resp.totals == totals_local[:len(resp.totals)], "e2e path and resident path disagree"
    e2e = {"value": world * n * C / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(resp.stats["h2d_bytes"]) * world,
           "d2h_bytes_per_step": int(resp.stats["d2h_bytes"] + 16 * C) * world, "steps": e2e_steps,
           "breakdown_ms": {k: round(resp.stats[k], 3) for k in ("flatten_ms", "h2d_ms", "kernel_ms", "d2h_ms")},
           "host_threads_per_gpu": host_threads}
    # (filled in below on rank 0) e2e["with_messages"]: the same path with every violation message rendered

    # the same call asked to also render every violation's {msg, details} (what Client.Review returns in the reference):
    # measured on a bounded sample, reported beside the decision-only figure
    msg_n = min(n, args.msg_sample)
    e2e_msgs = None
    if msg_n > 0 and rank == 0:
        sub = W.synth_objects(rank * n, msg_n, threads=host_threads)
        r2 = drv.ReviewBlob(sub, ep, flags=D.F_MATERIALIZE, with_results=False)   # results stay in the engine's buffers: not converted to Python
        dt = sum(r2.stats[k] for k in ("flatten_ms", "h2d_ms", "kernel_ms", "d2h_ms", "materialize_ms")) / 1e3
        e2e_msgs = {"value": msg_n * C / dt, "unit": UNIT, "objects": msg_n, "results_rendered": r2.stats["n_violations"],
                    "materialize_ms": round(r2.stats["materialize_ms"], 1), "flatten_ms": round(r2.stats["flatten_ms"], 1)}
        del r2
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int (u32 ids / u8 bytes / i64)",
        "data": "synthetic",
        "config": {"workload": "audit sweep: 10 gatekeeper in-tree templates, 50 constraints, 1M synthetic Pods per GPU (BASELINE.json configs[1])",
                   "objects_per_gpu": n, "constraints": C, "evals_per_step": world * n * C,
                   "l2": ("inputs larger than L2: %.0f MB of columns per GPU" % (alg_in / 1e6)) if flush is None else "explicit 256 MB L2 flush between steps",
                   "parallelism": (f"objects sharded over {world} GPU(s); exchange per step: " +
                                   ("fused into the kernel -- bitmap words and totals stored straight into every peer's buffer over NVLink, one device-side barrier"
                                    if sweep.p2p is not None else "one NCCL all-gather (bitmap shard + totals)")) if world > 1 else "single GPU",
                   "violating_pairs_per_step": int(sum(totals_host)) * 1, "synth_s": round(gen_s, 2),
                   "flatten_ms_once": round(rb.stats["flatten_ms"], 1)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "gk_eval_kernel", "kernel_ms": kern_ms, "alg_bytes_per_launch": alg_bytes, "peak_source": peak_src},
        "e2e": dict(e2e, with_messages=e2e_msgs), "gpu_launches": int(launches) if launches > 0 else K, "clocks": clocks,
    }
    if world == 1:
        line["cpu_baseline"] = cpu_baseline_single(args.cpu_sample)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--objects", type=int, default=1_000_000, help="objects per GPU (weak scaling)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=2000, help="Pods reviewed by the single-core oracle leg (about 11 s of CPU)")
    ap.add_argument("--msg-sample", type=int, default=50_000, help="objects of the sample whose messages are all rendered (e2e.with_messages)")
    ap.add_argument("--ref-objects-per-core", type=int, default=48)
    args = ap.parse_args()
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()
