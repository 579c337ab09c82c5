"""Latency / throughput definitions of the reference's own benchmark harness (`gator bench`), so that the numbers this
repo reports for the admission replay (BASELINE.json configs[2]) mean what the reference's mean.

Durations are integer nanoseconds (Go's time.Duration).
"""
from __future__ import annotations

from typing import Dict, Sequence


def percentile(sorted_ns: Sequence[int], p: float) -> int:
    """percentile -- pkg/gator/bench/metrics.go:37-59: linear interpolation at rank p/100 * (n-1) on an ascending
    slice; 0 for an empty slice."""
    n = len(sorted_ns)
    if n == 0:
        return 0
    if n == 1:
        return int(sorted_ns[0])
    rank = (p / 100.0) * float(n - 1)
    lower = int(rank)
    upper = lower + 1
    if upper >= n:
        return int(sorted_ns[-1])
    weight = rank - float(lower)
    return int(float(sorted_ns[lower]) * (1 - weight) + float(sorted_ns[upper]) * weight)


def calculate_latencies(durations_ns: Sequence[int]) -> Dict[str, int]:
    """calculateLatencies -- pkg/gator/bench/metrics.go:9-35."""
    if not durations_ns:
        return {"min": 0, "max": 0, "mean": 0, "p50": 0, "p95": 0, "p99": 0}
    s = sorted(int(d) for d in durations_ns)
    total = sum(s)
    q, r = divmod(abs(total), len(s))            # Go integer division truncates toward zero
    mean = q if total >= 0 else -q
    return {"min": s[0], "max": s[-1], "mean": mean, "p50": percentile(s, 50), "p95": percentile(s, 95), "p99": percentile(s, 99)}


def calculate_throughput(review_count: int, duration_ns: int) -> float:
    """calculateThroughput -- pkg/gator/bench/metrics.go:61-67: reviews per second, 0 for a zero duration."""
    if duration_ns == 0:
        return 0.0
    return float(review_count) / (duration_ns / 1e9)
