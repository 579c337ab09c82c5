"""How many host cores does this box really give us?  (cgroup quota, affinity, and measured scaling of the flattener
and of a pure-Python busy loop.)  Usage: python tools/host_probe.py [n_objects]"""
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _spin(_):
    t0 = time.perf_counter()
    x = 0
    for i in range(3_000_000):
        x += i * i
    return time.perf_counter() - t0


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us", "/proc/loadavg"):
        try:
            print(p, open(p).read().strip())
        except OSError:
            pass
    try:
        model = [l for l in open("/proc/cpuinfo") if l.startswith("model name")]
        print("cpu:", model[0].split(":", 1)[1].strip(), "x", len(model))
    except OSError:
        pass
    single = None
    for procs in (1, 8, 32, 64, 128):
        if procs > (os.cpu_count() or 1):
            break
        with mp.Pool(procs) as pool:
            t0 = time.perf_counter()
            each = pool.map(_spin, range(procs))
            wall = time.perf_counter() - t0
        if single is None:
            single = each[0]
        print("busy loop: %3d procs  wall %.2fs  per-proc mean %.2fs  -> throughput x%.1f of one core" % (procs, wall, sum(each) / procs, procs * single / wall), flush=True)
    from gatekeeper_b200 import driver as D
    from gatekeeper_b200 import workloads as W
    lib = os.environ.get("GK_PROBE_LIB")
    blob = W.synth_objects(0, n)
    tm, cons = W.config2()
    for th in (1, 8, 16, 32, 64, 128, 256):
        if th > 2 * (os.cpu_count() or 1):
            break
        drv = D.Driver(lib_path=lib, threads=th)
        for k, r in tm:
            drv.add_template(k, r)
        for c in cons:
            drv.AddConstraint(c)
        for ns in W.synth_namespaces():
            drv.AddData("t", ["cluster", "v1", "Namespace", ns["metadata"]["name"]], ns)
        t0 = time.perf_counter()
        rb = drv.upload_blob(blob)
        dt = time.perf_counter() - t0
        print("flatten: threads %3d  %.2f s  %.2f us/obj  (cpu-us/obj %.1f)" % (th, dt, 1e6 * dt / n, 1e6 * dt / n * th), flush=True)
        rb.free()
        drv.close()


if __name__ == "__main__":
    main()
