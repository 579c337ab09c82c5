// CUDA backend for sm_100a: one thread per object evaluates every constraint (match pre-filter + lowered
// predicate) against the column-wise batch resident in HBM.
//
//   * the constraint table (match blocks, instructions, constant pools) is staged into shared memory once
//     per CTA; every lane of a warp walks the same constraint at the same time, so table reads broadcast;
//   * header columns are read with unit stride across the warp (coalesced 128 B lines); scope/label CSR
//     rows of neighbouring objects are adjacent in memory, so the ragged reads stay within a few lines;
//   * per-constraint totals use warp ballot + popc into shared counters, one global atomic per CTA;
//   * each thread assembles its own 32-constraint bitmap words in registers and stores them once.
// This is integer / byte work bounded by HBM traffic -- there is nothing to put on tensor cores.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "backend.hpp"
#include "vm_core.h"

namespace gk {

#define CK(x)                                                                                         \
  do {                                                                                                \
    cudaError_t e_ = (x);                                                                             \
    if (e_ != cudaSuccess) throw BackendError{std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #x}; \
  } while (0)

struct KParams {
  GkBatch batch;
  GkProgram prog;
  GkOut out;
  const uint32_t* active;     // [nconstraints] enforcement-point filter
  const uint32_t* slot_off;   // [nslots] word offset of each slot inside the slot area (depends on the batch's tile capacities)
  const uint32_t* tile_lo;    // [(ntiles + 1) * nscopes] first row of every scope for every tile (row ranges are contiguous)
  uint32_t ntiles;
  uint32_t tile;              // objects per tile (multiple of 32)
  uint32_t slot_words;        // words in the slot area
};

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr uint32_t kTile = 512;

__device__ __forceinline__ void stage(void* dst, const void* src, size_t bytes) {
  // 16-byte vector copies; sizes/offsets are padded to 16 on the host
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (size_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = s[i];
}

__device__ __forceinline__ uint32_t range_mask(uint32_t w, uint32_t a, uint32_t b) {
  // bits of word w that lie in the row range [a, b)
  const uint32_t lo = w * 32u, hi = lo + 32u;
  const uint32_t x = a > lo ? a : lo, y = b < hi ? b : hi;
  if (x >= y) return 0u;
  const uint32_t nb = y - x;
  return (nb == 32u ? 0xffffffffu : ((1u << nb) - 1u)) << (x - lo);
}

// One CTA = one tile of consecutive objects at a time; all intermediate bit columns live in shared memory.
__global__ void __launch_bounds__(kThreads) gk_eval_kernel(const KParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t C = p.prog.nconstraints, W = p.out.words, NS = p.batch.nscopes;
  const uint32_t TW = p.tile / 32u;
  // ---- shared-memory layout
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* q = smem + off;
    off += (bytes + 15) / 16 * 16;
    return q;
  };
  uint32_t* s_tot = static_cast<uint32_t*>(take((size_t)C * 4));
  uint32_t* s_err = static_cast<uint32_t*>(take((size_t)C * 4));
  uint32_t* s_act = static_cast<uint32_t*>(take((size_t)C * 4));
  uint32_t* s_lo = static_cast<uint32_t*>(take((size_t)NS * 4));
  uint32_t* s_cnt = static_cast<uint32_t*>(take((size_t)NS * 4));
  uint32_t* s_soff = static_cast<uint32_t*>(take((size_t)p.prog.nslots * 4));
  uint32_t* res_v = static_cast<uint32_t*>(take((size_t)p.tile * W * 4));
  uint32_t* res_e = static_cast<uint32_t*>(take((size_t)p.tile * W * 4));
  uint32_t* slots = static_cast<uint32_t*>(take((size_t)p.slot_words * 4));
  GkOp* ops = static_cast<GkOp*>(take((size_t)p.prog.nops * sizeof(GkOp)));
  GkMatch* match = static_cast<GkMatch*>(take((size_t)p.prog.nmatch * sizeof(GkMatch)));
  GkColumn* cols = static_cast<GkColumn*>(take((size_t)p.batch.ncols * sizeof(GkColumn)));
  GkScope* scopes = static_cast<GkScope*>(take((size_t)NS * sizeof(GkScope)));
  uint32_t* pool = static_cast<uint32_t*>(take((size_t)p.prog.npool * 4));
  uint8_t* cbytes = static_cast<uint8_t*>(take((size_t)p.prog.ncbytes));

  for (uint32_t i = threadIdx.x; i < C; i += blockDim.x) {
    s_tot[i] = 0;
    s_err[i] = 0;
    s_act[i] = p.active[i];
  }
  for (uint32_t i = threadIdx.x; i < p.prog.nslots; i += blockDim.x) s_soff[i] = p.slot_off[i];
  stage(ops, p.prog.ops, ((size_t)p.prog.nops * sizeof(GkOp) + 15) / 16 * 16);
  stage(match, p.prog.match, ((size_t)p.prog.nmatch * sizeof(GkMatch) + 15) / 16 * 16);
  stage(cols, p.batch.cols, ((size_t)p.batch.ncols * sizeof(GkColumn) + 15) / 16 * 16);
  stage(scopes, p.batch.scopes, ((size_t)NS * sizeof(GkScope) + 15) / 16 * 16);
  stage(pool, p.prog.pool, ((size_t)p.prog.npool * 4 + 15) / 16 * 16);
  stage(cbytes, p.prog.cbytes, ((size_t)p.prog.ncbytes + 15) / 16 * 16);
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t FULL = 0xffffffffu;

  for (uint32_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
    // ---- tile row ranges (precomputed on the host: rows of a tile are contiguous at every scope)
    for (uint32_t s = threadIdx.x; s < NS; s += blockDim.x) {
      const uint32_t a = p.tile_lo[(size_t)t * NS + s], b = p.tile_lo[(size_t)(t + 1) * NS + s];
      s_lo[s] = a;
      s_cnt[s] = b - a;
    }
    for (uint32_t i = threadIdx.x; i < p.tile * W; i += blockDim.x) {
      res_v[i] = 0;
      res_e[i] = 0;
    }
    __syncthreads();
    const uint32_t nobj = s_cnt[0], obj0 = s_lo[0];

    uint32_t pc = 0;
    for (;;) {   // phases
      uint32_t k = 0;
      bool end = false;
      for (;; ++pc, ++k) {
        const GkOp op = ops[pc];
        const uint32_t kind = op.w0 & 0xffu;
        if (kind == GK_N_PHASE) {
          ++pc;
          break;
        }
        if (kind == GK_N_END) {
          end = true;
          break;
        }
        if (k % kWarps != warp) continue;   // ops of a phase are independent: one warp per op, round-robin
        const uint32_t level = (op.w0 >> 8) & 0xffu;
        uint32_t* out = slots + s_soff[op.w0 >> 16];
        switch (kind) {
          case GK_N_ATOM: {
            const GkColumn& c = cols[op.w1 >> 8];
            const uint32_t aop = op.w1 & 0xffu, lo = s_lo[level], cnt = s_cnt[level];
            for (uint32_t r = lane; r < ((cnt + 31u) & ~31u); r += 32u) {
              const bool v = r < cnt && gk_atom(c, lo + r, aop, op.w2, op.w3, pool, cbytes);
              const uint32_t w = __ballot_sync(FULL, v);
              if (lane == 0) out[r >> 5] = w;
            }
            break;
          }
          case GK_N_GATE: {
            const uint32_t* a = slots + s_soff[op.w1 & 0xffffu];
            const uint32_t* b = slots + s_soff[op.w1 >> 16];
            const uint32_t f = op.w2, words = (s_cnt[level] + 31u) >> 5;
            const uint32_t na = (f & GK_G_NEG_A) ? FULL : 0u, nb = (f & GK_G_NEG_B) ? FULL : 0u, no = (f & GK_G_NEG_OUT) ? FULL : 0u;
            for (uint32_t i = lane; i < words; i += 32u) {
              const uint32_t x = a[i] ^ na, y = b[i] ^ nb;
              out[i] = ((f & GK_G_OR) ? (x | y) : (x & y)) ^ no;
            }
            break;
          }
          case GK_N_CONST: {
            const uint32_t v = (op.w1 & 1u) ? FULL : 0u, words = (s_cnt[level] + 31u) >> 5;
            for (uint32_t i = lane; i < words; i += 32u) out[i] = v;
            break;
          }
          case GK_N_BCAST: {   // parent-level column -> rows of the child scope `level`
            const uint32_t* in = slots + s_soff[op.w1 & 0xffffu];
            const uint32_t par = (uint32_t)scopes[level].parent;
            const uint32_t* coff = scopes[level].off + s_lo[par];
            const uint32_t clo = s_lo[level], pcnt = s_cnt[par], words = (s_cnt[level] + 31u) >> 5;
            for (uint32_t i = lane; i < words; i += 32u) out[i] = 0u;
            __syncwarp();
            for (uint32_t r = lane; r < pcnt; r += 32u) {
              if ((in[r >> 5] >> (r & 31u)) & 1u) {
                const uint32_t a = coff[r] - clo, b = coff[r + 1] - clo;
                if (b > a)
                  for (uint32_t w = a >> 5; w <= (b - 1u) >> 5; ++w) atomicOr(&out[w], range_mask(w, a, b));
              }
            }
            break;
          }
          case GK_N_ACC: {     // EXISTS: OR over each parent's child range of the scope `level`
            const uint32_t* in = slots + s_soff[op.w1 & 0xffffu];
            const uint32_t par = (uint32_t)scopes[level].parent;
            const uint32_t* coff = scopes[level].off + s_lo[par];
            const uint32_t clo = s_lo[level], pcnt = s_cnt[par];
            for (uint32_t r = lane; r < ((pcnt + 31u) & ~31u); r += 32u) {
              bool any = false;
              if (r < pcnt) {
                const uint32_t a = coff[r] - clo, b = coff[r + 1] - clo;
                if (b > a)
                  for (uint32_t w = a >> 5; w <= (b - 1u) >> 5 && !any; ++w) any = (in[w] & range_mask(w, a, b)) != 0u;
              }
              const uint32_t w = __ballot_sync(FULL, any);
              if (lane == 0) out[r >> 5] = w;
            }
            break;
          }
          case GK_N_MATCH: {
            uint32_t* err = slots + s_soff[op.w1 & 0xffffu];
            const GkMatch& m = match[op.w2];
            for (uint32_t r = lane; r < ((nobj + 31u) & ~31u); r += 32u) {
              int res = 0;
              if (r < nobj && !(p.batch.flags[obj0 + r] & GK_F_SKIP)) res = gk_match(p.batch, pool, cbytes, m, obj0 + r);
              if (res < 0) {
                const uint32_t slot = atomicAdd(p.out.errcount, 1u);
                if (slot < p.out.errcap) {
                  p.out.errlist[3 * slot] = obj0 + r;
                  p.out.errlist[3 * slot + 1] = op.w2;
                  p.out.errlist[3 * slot + 2] = (uint32_t)(-res);
                }
              }
              const uint32_t wm = __ballot_sync(FULL, res > 0), we = __ballot_sync(FULL, res < 0);
              if (lane == 0) {
                out[r >> 5] = wm;
                err[r >> 5] = we;
              }
            }
            break;
          }
          case GK_N_OUT: {
            const uint32_t c = op.w2, flags = op.w3 >> 16;
            if (!s_act[c]) break;
            const uint32_t* prog = slots + s_soff[op.w1 & 0xffffu];
            const uint32_t* mt = slots + s_soff[op.w1 >> 16];
            const uint32_t* er = slots + s_soff[op.w3 & 0xffffu];
            const uint32_t bit = 1u << (c & 31u), wi = c >> 5;
            uint32_t nv = 0, ne = 0;
            for (uint32_t i = 0; i < ((nobj + 31u) >> 5); ++i) {
              const uint32_t valid = range_mask(i, 0u, nobj);
              const uint32_t pv = (flags & 1u) ? FULL : (flags & 2u) ? 0u : prog[i];
              const uint32_t v = pv & mt[i] & valid, e = er[i] & valid;
              const uint32_t o = i * 32u + lane;
              if ((v >> lane) & 1u) atomicOr(&res_v[o * W + wi], bit);
              if ((e >> lane) & 1u) atomicOr(&res_e[o * W + wi], bit);
              nv += __popc(v);
              ne += __popc(e);
            }
            if (lane == 0) {
              if (nv) atomicAdd(&s_tot[c], nv);
              if (ne) atomicAdd(&s_err[c], ne);
            }
            break;
          }
          default: break;
        }
      }
      __syncthreads();
      if (end) break;
    }
    // ---- the tile's bitmap rows are contiguous in the object-major output: coalesced copy out
    const size_t base = (size_t)obj0 * W;
    for (uint32_t i = threadIdx.x; i < nobj * W; i += blockDim.x) {
      p.out.viol[base + i] = res_v[i];
      p.out.err[base + i] = res_e[i];
    }
    __syncthreads();
  }
  for (uint32_t c = threadIdx.x; c < C; c += blockDim.x) {
    if (s_tot[c]) atomicAdd(p.out.totals + c, (unsigned long long)s_tot[c]);
    if (s_err[c]) atomicAdd(p.out.err_totals + c, (unsigned long long)s_err[c]);
  }
}

// ------------------------------------------------------------------------------------------------ backend
struct DevBatch {
  uint8_t* arena = nullptr;
  size_t bytes = 0;
  GkBatch hdr{};
  uint32_t n = 0;
  // per-batch output buffers (reused by every eval of a resident batch)
  uint32_t* viol = nullptr;
  uint32_t* err = nullptr;
  uint32_t words = 0;
  // tiling (depends on the batch's row distribution and on the program's slot table)
  uint32_t* d_tile_lo = nullptr;
  uint32_t* d_slot_off = nullptr;
  uint32_t ntiles = 0, slot_words = 0;
  uint64_t prog_version = 0;
};

class CudaBackend : public Backend {
 public:
  explicit CudaBackend(int device) : device_(device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw BackendError{"no CUDA device: gatekeeper_b200 has no CPU fallback; the evaluation path requires a B200 (sm_100a) GPU"};
    if (device_ >= ndev) throw BackendError{"CUDA device index out of range"};
    CK(cudaSetDevice(device_));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device_));
    sms_ = prop.multiProcessorCount;
    max_smem_ = (size_t)prop.sharedMemPerBlockOptin;
    sm_smem_ = (size_t)prop.sharedMemPerMultiprocessor;
    CK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    CK(cudaEventCreate(&ev0_));
    CK(cudaEventCreate(&ev1_));
    CK(cudaFuncSetAttribute(gk_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem_));
    CK(cudaMalloc(&d_scalars_, 4096));
  }
  ~CudaBackend() override {
    cudaSetDevice(device_);
    free_tables();
    if (d_dict_off_) cudaFree(d_dict_off_);
    if (d_dict_bytes_) cudaFree(d_dict_bytes_);
    if (d_scalars_) cudaFree(d_scalars_);
    if (d_totals_) cudaFree(d_totals_);
    if (d_errlist_) cudaFree(d_errlist_);
    if (d_active_) cudaFree(d_active_);
    if (pinned_) cudaFreeHost(pinned_);
    cudaEventDestroy(ev0_);
    cudaEventDestroy(ev1_);
    cudaStreamDestroy(stream_);
  }
  const char* name() const override { return "cuda-sm100a"; }

  void set_program(const Compiled& c) override {
    std::lock_guard<std::mutex> l(mu_);
    if (version_ == c.version) return;
    CK(cudaSetDevice(device_));
    free_tables();
    prog_ = GkProgram{};
    prog_.nconstraints = (uint32_t)c.cons_match.size();
    prog_.nmatch = (uint32_t)c.match.size();
    prog_.nops = (uint32_t)c.ops.size();
    prog_.nslots = (uint32_t)c.slot_level.size();
    prog_.npool = (uint32_t)c.pool.size();
    prog_.ncbytes = (uint32_t)c.cbytes.size();
    auto up = [&](const void* src, size_t bytes, void** dst) {
      size_t padded = (bytes + 63) / 64 * 64 + 64;
      CK(cudaMalloc(dst, padded));
      CK(cudaMemset(*dst, 0, padded));
      if (bytes) CK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    };
    up(c.ops.data(), c.ops.size() * sizeof(GkOp), (void**)&d_ops_);
    up(c.match.data(), c.match.size() * sizeof(GkMatch), (void**)&d_match_);
    up(c.pool.data(), c.pool.size() * 4, (void**)&d_pool_);
    up(c.cbytes.data(), c.cbytes.size(), (void**)&d_cbytes_);
    prog_.ops = d_ops_;
    prog_.match = d_match_;
    prog_.pool = d_pool_;
    prog_.cbytes = d_cbytes_;
    slot_level_ = c.slot_level;
    scope_parent_.clear();
    for (auto& sd : c.schema.scopes) scope_parent_.push_back(sd.parent);
    ncols_ = (uint32_t)c.schema.cols.size();
    const uint32_t C = prog_.nconstraints;
    if (d_totals_) cudaFree(d_totals_);
    CK(cudaMalloc(&d_totals_, (size_t)(2 * std::max(C, 1u)) * sizeof(unsigned long long)));
    if (d_active_) cudaFree(d_active_);
    CK(cudaMalloc(&d_active_, (size_t)std::max(C, 1u) * 4));
    if (!d_errlist_) CK(cudaMalloc(&d_errlist_, (size_t)kErrCap * 3 * 4));
    last_active_.clear();
    version_ = c.version;
  }

  void sync_strings(const StringTable& st) override {
    std::lock_guard<std::mutex> l(mu_);
    uint32_t n = st.size();
    if (n == dict_n_) return;
    CK(cudaSetDevice(device_));
    std::vector<uint32_t> off;
    std::vector<uint8_t> bytes;
    st.snapshot(off, bytes);
    if (d_dict_off_) cudaFree(d_dict_off_);
    if (d_dict_bytes_) cudaFree(d_dict_bytes_);
    CK(cudaMalloc(&d_dict_off_, off.size() * 4 + 64));
    CK(cudaMalloc(&d_dict_bytes_, bytes.size() + 64));
    CK(cudaMemcpy(d_dict_off_, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    if (!bytes.empty()) CK(cudaMemcpy(d_dict_bytes_, bytes.data(), bytes.size(), cudaMemcpyHostToDevice));
    dict_n_ = (uint32_t)off.size() - 1;
  }

  void* upload(const HostBatch& hb, const Compiled& c, double* h2d_ms, uint64_t* h2d_bytes) override {
    CK(cudaSetDevice(device_));
    PackedBatch pb;
    pack_batch(hb, c, pb);
    // ---- tiling: first row of every scope for every tile, slot offsets from the per-scope tile capacities
    const uint32_t NS = (uint32_t)c.schema.scopes.size();
    const uint32_t ntiles = (hb.n + kTile - 1) / kTile;
    std::vector<uint32_t> tile_lo((size_t)(ntiles + 1) * NS), cap(NS, 0);
    for (uint32_t t = 0; t <= ntiles; ++t) {
      uint32_t* lo = &tile_lo[(size_t)t * NS];
      lo[0] = std::min<uint32_t>(t * kTile, hb.n);
      for (uint32_t s = 1; s < NS; ++s) lo[s] = hb.scope_off[s][lo[c.schema.scopes[s].parent]];
      if (t)
        for (uint32_t s = 0; s < NS; ++s) cap[s] = std::max(cap[s], lo[s] - tile_lo[(size_t)(t - 1) * NS + s]);
    }
    cap[0] = kTile;
    std::vector<uint32_t> slot_off(c.slot_level.size());
    uint32_t slot_words = 0;
    for (size_t i = 0; i < slot_off.size(); ++i) {
      slot_off[i] = slot_words;
      slot_words += (cap[c.slot_level[i]] + 31) / 32 + 1;
    }
    auto* db = new DevBatch();
    db->bytes = gk_align(pb.arena.size());
    db->n = hb.n;
    db->ntiles = ntiles;
    db->slot_words = slot_words;
    db->prog_version = c.version;
    CK(cudaMalloc(&db->arena, db->bytes));
    CK(cudaMalloc(&db->d_tile_lo, tile_lo.size() * 4 + 64));
    CK(cudaMalloc(&db->d_slot_off, slot_off.size() * 4 + 64));
    db->hdr = rebase_batch(pb, pb.arena.data(), db->arena);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    {
      std::lock_guard<std::mutex> l(mu_);
      // stage through pinned memory so the copy runs at full PCIe rate
      if (pinned_bytes_ < pb.arena.size()) {
        if (pinned_) cudaFreeHost(pinned_);
        pinned_bytes_ = pb.arena.size() + (pb.arena.size() >> 2);
        CK(cudaMallocHost(&pinned_, pinned_bytes_));
      }
      memcpy(pinned_, pb.arena.data(), pb.arena.size());
      CK(cudaEventRecord(a, stream_));
      CK(cudaMemcpyAsync(db->arena, pinned_, pb.arena.size(), cudaMemcpyHostToDevice, stream_));
      CK(cudaMemcpyAsync(db->d_tile_lo, tile_lo.data(), tile_lo.size() * 4, cudaMemcpyHostToDevice, stream_));
      CK(cudaMemcpyAsync(db->d_slot_off, slot_off.data(), slot_off.size() * 4, cudaMemcpyHostToDevice, stream_));
      CK(cudaEventRecord(b, stream_));
      CK(cudaStreamSynchronize(stream_));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (h2d_ms) *h2d_ms = ms;
    if (h2d_bytes) *h2d_bytes = pb.arena.size() + tile_lo.size() * 4 + slot_off.size() * 4;
    db->words = std::max<uint32_t>(1, (uint32_t)((c.cons_match.size() + 31) / 32));
    CK(cudaMalloc(&db->viol, (size_t)std::max(db->n, 1u) * db->words * 4));
    CK(cudaMalloc(&db->err, (size_t)std::max(db->n, 1u) * db->words * 4));
    return db;
  }

  void release(void* b) override {
    auto* db = static_cast<DevBatch*>(b);
    if (!db) return;
    cudaSetDevice(device_);
    cudaFree(db->arena);
    cudaFree(db->viol);
    cudaFree(db->err);
    cudaFree(db->d_tile_lo);
    cudaFree(db->d_slot_off);
    delete db;
  }

  // everything a launch needs besides the kernel itself (kept outside the timed region)
  KParams prepare(DevBatch* db, const std::vector<uint32_t>& active, uint32_t* viol, uint32_t* err, unsigned long long* totals,
                  unsigned long long* err_totals, cudaStream_t st, size_t* smem_out) {
    const uint32_t C = prog_.nconstraints;
    if (db->prog_version != version_) throw BackendError{"batch was uploaded for another constraint set"};
    KParams p;
    p.batch = db->hdr;
    p.batch.dict_off = d_dict_off_;
    p.batch.dict_bytes = d_dict_bytes_;
    p.batch.dict_n = dict_n_;
    p.prog = prog_;
    p.out.viol = viol;
    p.out.err = err;
    p.out.totals = totals;
    p.out.err_totals = err_totals;
    p.out.errlist = d_errlist_;
    p.out.errcount = reinterpret_cast<uint32_t*>(d_scalars_);
    p.out.errcap = kErrCap;
    p.out.words = db->words;
    p.active = d_active_;
    p.slot_off = db->d_slot_off;
    p.tile_lo = db->d_tile_lo;
    p.ntiles = db->ntiles;
    p.tile = kTile;
    p.slot_words = db->slot_words;
    if (active.size() != C) throw BackendError{"active mask size mismatch"};
    if (C && active != last_active_) {
      CK(cudaMemcpyAsync(d_active_, active.data(), (size_t)C * 4, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));   // `active` is a caller temporary
      last_active_ = active;
    }
    if (C) CK(cudaMemsetAsync(totals, 0, (size_t)C * sizeof(unsigned long long), st));
    if (C) CK(cudaMemsetAsync(err_totals, 0, (size_t)C * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(d_scalars_, 0, 64, st));
    auto r16 = [](size_t x) { return (x + 15) / 16 * 16; };
    const size_t NS = scope_parent_.size();
    size_t smem = 3 * r16((size_t)C * 4) + 2 * r16(NS * 4) + r16((size_t)prog_.nslots * 4) + 2 * r16((size_t)kTile * db->words * 4) +
                  r16((size_t)db->slot_words * 4) + r16((size_t)prog_.nops * sizeof(GkOp)) + r16((size_t)prog_.nmatch * sizeof(GkMatch)) +
                  r16((size_t)ncols_ * sizeof(GkColumn)) + r16(NS * sizeof(GkScope)) + r16((size_t)prog_.npool * 4) + r16((size_t)prog_.ncbytes) + 64;
    if (smem > max_smem_)
      throw BackendError{"constraint set needs " + std::to_string(smem) + " bytes of shared memory per CTA (limit " + std::to_string(max_smem_) +
                         "): too many live netlist columns for one launch"};
    *smem_out = smem;
    return p;
  }

  void fire(const KParams& p, size_t smem, cudaStream_t st) {
    if (p.ntiles == 0) return;
    // persistent CTAs: as many as fit per SM (shared-memory bound), each walks tiles with a grid stride
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / kThreads, (sm_smem_ - 1024) / (smem + 1024)));
    uint32_t grid = std::max(1u, std::min<uint32_t>(p.ntiles, (uint32_t)(sms_ * per_sm)));
    gk_eval_kernel<<<grid, kThreads, smem, st>>>(p);
    CK(cudaGetLastError());
    ++launches_;
  }

  void eval(void* b, const std::vector<uint32_t>& active, EvalOut& out, bool copy_back) override {
    auto* db = static_cast<DevBatch*>(b);
    std::lock_guard<std::mutex> l(mu_);
    CK(cudaSetDevice(device_));
    const uint32_t C = prog_.nconstraints;
    out.n = db->n;
    out.nconstraints = C;
    out.words = db->words;
    unsigned long long* totals = d_totals_;
    unsigned long long* err_totals = d_totals_ + std::max(C, 1u);
    size_t smem = 0;
    KParams p = prepare(db, active, db->viol, db->err, totals, err_totals, stream_, &smem);
    CK(cudaStreamSynchronize(stream_));
    // the event pair brackets exactly the evaluation kernel
    CK(cudaEventRecord(ev0_, stream_));
    fire(p, smem, stream_);
    CK(cudaEventRecord(ev1_, stream_));
    CK(cudaStreamSynchronize(stream_));
    CK(cudaEventElapsedTime(&out.kernel_ms, ev0_, ev1_));
    out.launches = launches_;
    out.totals.assign(C, 0);
    out.err_totals.assign(C, 0);
    if (C) {
      std::vector<unsigned long long> t(2 * (size_t)std::max(C, 1u));
      CK(cudaMemcpy(t.data(), d_totals_, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      for (uint32_t c = 0; c < C; ++c) {
        out.totals[c] = t[c];
        out.err_totals[c] = t[std::max(C, 1u) + c];
      }
    }
    uint32_t nerr = 0;
    CK(cudaMemcpy(&nerr, d_scalars_, 4, cudaMemcpyDeviceToHost));
    nerr = std::min(nerr, kErrCap);
    out.errlist.resize((size_t)nerr * 3);
    if (nerr) CK(cudaMemcpy(out.errlist.data(), d_errlist_, (size_t)nerr * 12, cudaMemcpyDeviceToHost));
    if (copy_back) {
      out.viol.resize((size_t)db->n * db->words);
      out.err.resize((size_t)db->n * db->words);
      if (db->n) {
        CK(cudaMemcpy(out.viol.data(), db->viol, out.viol.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.err.data(), db->err, out.err.size() * 4, cudaMemcpyDeviceToHost));
      }
    }
  }

  void eval_into(void* b, const std::vector<uint32_t>& active, const DevOutPtrs& dst) override {
    auto* db = static_cast<DevBatch*>(b);
    std::lock_guard<std::mutex> l(mu_);
    CK(cudaSetDevice(device_));
    size_t smem = 0;
    cudaStream_t st = static_cast<cudaStream_t>(dst.stream);
    KParams p = prepare(db, active, static_cast<uint32_t*>(dst.viol), static_cast<uint32_t*>(dst.err), static_cast<unsigned long long*>(dst.totals),
                        static_cast<unsigned long long*>(dst.err_totals), st, &smem);
    fire(p, smem, st);
  }

 private:
  static constexpr uint32_t kErrCap = 1u << 20;
  void free_tables() {
    if (d_ops_) cudaFree(d_ops_);
    if (d_match_) cudaFree(d_match_);
    if (d_pool_) cudaFree(d_pool_);
    if (d_cbytes_) cudaFree(d_cbytes_);
    d_ops_ = nullptr;
    d_match_ = nullptr;
    d_pool_ = nullptr;
    d_cbytes_ = nullptr;
  }
  int device_;
  int sms_ = 148;
  size_t max_smem_ = 0, sm_smem_ = 0;
  cudaStream_t stream_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  std::mutex mu_;
  uint64_t version_ = 0, launches_ = 0;
  GkProgram prog_{};
  std::vector<uint8_t> slot_level_;
  std::vector<int> scope_parent_;
  std::vector<uint32_t> last_active_;
  uint32_t ncols_ = 0;
  GkOp* d_ops_ = nullptr;
  GkMatch* d_match_ = nullptr;
  uint32_t* d_pool_ = nullptr;
  uint8_t* d_cbytes_ = nullptr;
  uint32_t* d_dict_off_ = nullptr;
  uint8_t* d_dict_bytes_ = nullptr;
  uint32_t dict_n_ = 0;
  void* d_scalars_ = nullptr;
  unsigned long long* d_totals_ = nullptr;
  uint32_t* d_errlist_ = nullptr;
  uint32_t* d_active_ = nullptr;
  void* pinned_ = nullptr;
  size_t pinned_bytes_ = 0;
};

Backend* make_backend(int device) { return new CudaBackend(device); }

}  // namespace gk
