// CUDA kernels of the device ingest path: one thread per object runs the shared per-object code of ingest_core.h.
//   gk_tape_kernel   raw JSON -> tape (+ review-level status)
//   gk_count_kernel  rows per scope / bytes per byte column / header byte counts of every object
//   gk_scan_kernel   exclusive prefix sums of the counter arrays (one CTA per array), totals
//   gk_header_kernel per object: flags, kind / group sids, names, labels, namespace row (the same steps in every thread)
//   gk_write_kernel  per object: the scope walk -- CSR scope offsets, row handles, byte-encoded columns
//   gk_cols_kernel   per row of one scope: every other column encoding; lookups through the device hash tables
//   gk_fill_kernel   the host's answers to the miss list -> table slots
//   gk_tiles_kernel  first row of every scope for every evaluation tile + the largest tile of each scope
// HBM-bound byte work (no tensor cores): the tape kernel reads the blob once (~1.2 KB per Pod) and writes ~1 tape entry per
// 8 bytes; count / write re-walk the tape, not the text.
#pragma once
#include <cuda_runtime.h>

#include "ingest_core.h"

namespace gk {

constexpr int kIngestThreads = 128;
#ifndef GK_INGEST_MIN_BLOCKS
#define GK_INGEST_MIN_BLOCKS 8   /* 64 registers per thread: these passes are latency-bound, resident warps are what hides it */
#endif

__global__ void __launch_bounds__(kIngestThreads) gk_tape_kernel(const GkIngestIn in, uint32_t first, uint32_t count) {
  const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < first + count) gk_tape_obj(in, i);
}

// count / write: one WARP per object (see gk_ingest_obj): the lanes share the object's tape through L1 and split its columns
constexpr uint32_t kMaxCounters = GK_MAX_SCOPES + 64 + GK_CNT_EXTRA;

// `lanes` (a power of two, 1..32) threads share one object
__global__ void __launch_bounds__(kIngestThreads) gk_count_kernel(const GkXProg xp, const GkIngestIn in, uint32_t lanes, uint32_t first, uint32_t count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, i = first + t / lanes, lane = t % lanes;
  if (i >= first + count) return;
  GkIngestOut none;
  memset(&none, 0, sizeof none);
  uint32_t cur[kMaxCounters];
  gk_ingest_obj<GK_PASS_COUNT>(xp, in, none, i, GkCur{cur, 1}, lane, lanes);
}

__global__ void __launch_bounds__(kIngestThreads, GK_INGEST_MIN_BLOCKS) gk_write_kernel(const GkXProg xp, const GkIngestIn in, const GkIngestOut out, uint32_t lanes) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, i = t / lanes, lane = t % lanes;
  if (i >= in.n) return;
  uint32_t cur[kMaxCounters];
  gk_ingest_obj<GK_PASS_ROWS>(xp, in, out, i, GkCur{cur, 1}, lane, lanes);
}

__global__ void __launch_bounds__(kIngestThreads) gk_header_kernel(const GkXProg xp, const GkIngestIn in, const GkIngestOut out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in.n) return;
  uint32_t cur[kMaxCounters];
  gk_ingest_obj<GK_PASS_HEADER>(xp, in, out, i, GkCur{cur, 1}, 0u, 1u);
}

// the per-row column pass of one scope: a thread (or `lanes` threads) per row
__global__ void __launch_bounds__(kIngestThreads, GK_INGEST_MIN_BLOCKS) gk_cols_kernel(const GkXProg xp, const GkIngestIn in, const GkIngestOut out, uint32_t scope, uint32_t rows,
                                                                 uint32_t lanes) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, r = t / lanes, lane = t % lanes;
  if (r >= rows) return;
  gk_ingest_row(xp, in, out, scope, r, lane, lanes);
}

// exclusive scan of counts[k * n .. (k + 1) * n) for k = blockIdx.x; totals[k] = the sum
__global__ void __launch_bounds__(1024) gk_scan_kernel(uint32_t* counts, uint32_t n, uint32_t* totals) {
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t carry;
  uint32_t* a = counts + (size_t)blockIdx.x * n;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 4096u) {
    // four consecutive elements per thread (one 16-byte load when aligned), block scan of the per-thread sums
    uint32_t v[4];
    uint32_t s = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t ix = base + threadIdx.x * 4u + u;
      v[u] = ix < n ? a[ix] : 0u;
      s += v[u];
    }
    uint32_t x = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if ((int)lane >= d) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
        if ((int)lane >= d) w += y;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    uint32_t excl = carry + (warp ? warp_sum[warp - 1] : 0u) + (x - s);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t ix = base + threadIdx.x * 4u + u;
      if (ix < n) a[ix] = excl;
      excl += v[u];
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_sum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

__global__ void gk_fill_kernel(uint32_t* vals, const uint32_t* pairs, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vals[pairs[2 * i]] = pairs[2 * i + 1];
}

// tile_lo[t * NS + s] = first row of scope s in tile t (t = ntiles: the totals); cap[s] = rows of the largest tile
__global__ void gk_tiles_kernel(const uint32_t* bases /* scanned counts */, const uint32_t* totals, uint32_t n, uint32_t NS, uint32_t tile, uint32_t ntiles,
                                uint32_t* tile_lo, uint32_t* cap) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  const uint32_t o = min(t * tile, n);
  for (uint32_t s = 0; s < NS; ++s) {
    const uint32_t lo = s == 0 ? o : (o < n ? bases[(size_t)s * n + o] : totals[s]);
    tile_lo[(size_t)t * NS + s] = lo;
    if (t < ntiles) {
      const uint32_t o2 = min((t + 1u) * tile, n);
      const uint32_t hi = s == 0 ? o2 : (o2 < n ? bases[(size_t)s * n + o2] : totals[s]);
      atomicMax(&cap[s], hi - lo);
    }
  }
}

}  // namespace gk
