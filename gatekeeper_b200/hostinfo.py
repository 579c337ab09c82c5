"""Host CPUs this process may really use: affinity mask and cgroup CPU quota respected (the same rule as
`gk_host_cpus()` in the native library; a container on a 128-thread host is often capped at a fraction of it)."""
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().split()
    except OSError:
        return None


def host_cpus() -> int:
    n = os.cpu_count() or 1
    try:
        n = min(n, max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        pass
    v2 = _read("/sys/fs/cgroup/cpu.max")
    if v2 and v2[0] != "max" and len(v2) >= 2 and float(v2[1]) > 0:
        n = min(n, max(1, int(float(v2[0]) / float(v2[1]) + 0.5)))
    q, p = _read("/sys/fs/cgroup/cpu/cpu.cfs_quota_us"), _read("/sys/fs/cgroup/cpu/cpu.cfs_period_us")
    if q and p and float(q[0]) > 0 and float(p[0]) > 0:
        n = min(n, max(1, int(float(q[0]) / float(p[0]) + 0.5)))
    return n
