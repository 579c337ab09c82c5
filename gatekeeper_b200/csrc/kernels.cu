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
  const uint32_t* active;   // [nconstraints]
  uint32_t smem_tables;     // 1: tables staged in shared memory
};

constexpr int kThreads = 128;

__device__ __forceinline__ void stage(void* dst, const void* src, size_t bytes) {
  // 16-byte vector copies; sizes/offsets are padded to 16 on the host
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (size_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = s[i];
}

__global__ void __launch_bounds__(kThreads) gk_eval_kernel(const KParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t C = p.prog.nconstraints;
  const uint32_t W = p.out.words;
  // ---- shared-memory layout: [totals C u32][err totals C u32][active C u32] then (when they fit) the tables:
  //      [cons][match][columns][scopes][instr][pool][cbytes]
  uint32_t* s_tot = reinterpret_cast<uint32_t*>(smem);
  uint32_t* s_err = s_tot + C;
  uint32_t* s_act = s_err + C;
  size_t off = ((size_t)3 * C * 4 + 15) / 16 * 16;
  const GkCons* cons = p.prog.cons;
  const GkMatch* match = p.prog.match;
  const GkInstr* instr = p.prog.instr;
  const uint32_t* pool = p.prog.pool;
  const uint8_t* cbytes = p.prog.cbytes;
  const GkColumn* cols = p.batch.cols;
  const GkScope* scopes = p.batch.scopes;
  for (uint32_t i = threadIdx.x; i < 2 * C; i += blockDim.x) s_tot[i] = 0;
  for (uint32_t i = threadIdx.x; i < C; i += blockDim.x) s_act[i] = p.active[i];
  if (p.smem_tables) {
    auto place = [&](const void* src, size_t bytes) {
      void* dst = smem + off;
      stage(dst, src, (bytes + 15) / 16 * 16);
      off += (bytes + 15) / 16 * 16;
      return dst;
    };
    cons = static_cast<const GkCons*>(place(p.prog.cons, (size_t)C * sizeof(GkCons)));
    match = static_cast<const GkMatch*>(place(p.prog.match, (size_t)p.prog.nmatch * sizeof(GkMatch)));
    cols = static_cast<const GkColumn*>(place(p.batch.cols, (size_t)p.batch.ncols * sizeof(GkColumn)));
    scopes = static_cast<const GkScope*>(place(p.batch.scopes, (size_t)p.batch.nscopes * sizeof(GkScope)));
    instr = static_cast<const GkInstr*>(place(p.prog.instr, (size_t)p.prog.ninstr * sizeof(GkInstr)));
    pool = static_cast<const uint32_t*>(place(p.prog.pool, (size_t)p.prog.npool * 4));
    cbytes = static_cast<const uint8_t*>(place(p.prog.cbytes, (size_t)p.prog.ncbytes));
  }
  __syncthreads();

  const uint32_t n = p.batch.n;
  const uint32_t lane = threadIdx.x & 31u;
  // one object per thread, 32 consecutive objects per warp; every lane of a warp walks the same constraint and the
  // same instruction, so table reads broadcast and there is no divergent dispatch
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const uint32_t obj_raw = base + threadIdx.x;
    const uint32_t obj = obj_raw < n ? obj_raw : n - 1;
    const bool live = obj_raw < n && !(p.batch.flags[obj] & GK_F_SKIP);
    unsigned long long cse = 0, cse_valid = 0;
    uint32_t cur_mid = GK_NONE;
    int mres = 0;
    for (uint32_t w = 0; w < W; ++w) {
      uint32_t vbits = 0, ebits = 0;
      const uint32_t cend = min(C, (w + 1) * 32u);
      for (uint32_t c = w * 32u; c < cend; ++c) {
        if (!s_act[c]) continue;   // enforcement-point filter: warp-uniform
        const GkCons cc = cons[c];
        if (cc.match_id != cur_mid) {   // warp-uniform: constraints are grouped by match block
          cur_mid = cc.match_id;
          mres = live ? gk_match(p.batch, pool, cbytes, match[cur_mid], obj) : 0;
        }
        int flag = 0;
        bool v = cc.pc == GK_PC_ACCEPT ? true
                 : cc.pc == GK_PC_REJECT ? false
                                         : gk_eval_prog(cols, scopes, instr, pool, cbytes, cc.pc, obj, live, cse, cse_valid, &flag);
        v = v && mres > 0;
        int code = mres < 0 ? -mres : 0;
        if (mres > 0 && flag) {
          v = false;
          code = flag;
        }
        const bool e = code != 0;
        if (e) {
          const uint32_t slot = atomicAdd(p.out.errcount, 1u);
          if (slot < p.out.errcap) {
            p.out.errlist[3 * slot] = obj;
            p.out.errlist[3 * slot + 1] = c;
            p.out.errlist[3 * slot + 2] = (uint32_t)code;
          }
        }
        vbits |= (uint32_t)v << (c & 31u);
        ebits |= (uint32_t)e << (c & 31u);
        const uint32_t bv = __ballot_sync(0xffffffffu, v);
        const uint32_t be = __ballot_sync(0xffffffffu, e);
        if (lane == 0) {
          if (bv) atomicAdd(&s_tot[c], __popc(bv));
          if (be) atomicAdd(&s_err[c], __popc(be));
        }
      }
      if (obj_raw < n) {
        p.out.viol[(size_t)obj * W + w] = vbits;
        p.out.err[(size_t)obj * W + w] = ebits;
      }
    }
  }
  __syncthreads();
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
    prog_.nconstraints = (uint32_t)c.cons.size();
    prog_.nmatch = (uint32_t)c.match.size();
    prog_.ninstr = (uint32_t)c.instr.size();
    prog_.npool = (uint32_t)c.pool.size();
    prog_.ncbytes = (uint32_t)c.cbytes.size();
    auto up = [&](const void* src, size_t bytes, void** dst) {
      size_t padded = (bytes + 63) / 64 * 64 + 64;
      CK(cudaMalloc(dst, padded));
      CK(cudaMemset(*dst, 0, padded));
      if (bytes) CK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    };
    up(c.cons.data(), c.cons.size() * sizeof(GkCons), (void**)&d_cons_);
    up(c.match.data(), c.match.size() * sizeof(GkMatch), (void**)&d_match_);
    up(c.instr.data(), c.instr.size() * sizeof(GkInstr), (void**)&d_instr_);
    up(c.pool.data(), c.pool.size() * 4, (void**)&d_pool_);
    up(c.cbytes.data(), c.cbytes.size(), (void**)&d_cbytes_);
    prog_.cons = d_cons_;
    prog_.match = d_match_;
    prog_.instr = d_instr_;
    prog_.pool = d_pool_;
    prog_.cbytes = d_cbytes_;
    const uint32_t C = prog_.nconstraints;
    if (d_totals_) cudaFree(d_totals_);
    CK(cudaMalloc(&d_totals_, (size_t)(2 * std::max(C, 1u)) * sizeof(unsigned long long)));
    if (d_active_) cudaFree(d_active_);
    CK(cudaMalloc(&d_active_, (size_t)std::max(C, 1u) * 4));
    if (!d_errlist_) CK(cudaMalloc(&d_errlist_, (size_t)kErrCap * 3 * 4));
    auto r16 = [](size_t x) { return (x + 15) / 16 * 16; };
    smem_base_ = r16((size_t)3 * C * 4);
    smem_need_ = smem_base_ + r16((size_t)C * sizeof(GkCons)) + r16((size_t)prog_.nmatch * sizeof(GkMatch)) +
                 r16(c.schema.cols.size() * sizeof(GkColumn)) + r16(c.schema.scopes.size() * sizeof(GkScope)) +
                 r16((size_t)prog_.ninstr * sizeof(GkInstr)) + r16((size_t)prog_.npool * 4) + r16((size_t)prog_.ncbytes);
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
    auto* db = new DevBatch();
    db->bytes = gk_align(pb.arena.size());
    db->n = hb.n;
    CK(cudaMalloc(&db->arena, db->bytes));
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
      CK(cudaEventRecord(b, stream_));
      CK(cudaStreamSynchronize(stream_));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (h2d_ms) *h2d_ms = ms;
    if (h2d_bytes) *h2d_bytes = pb.arena.size();
    db->words = (uint32_t)((c.cons.size() + 31) / 32);
    if (db->words == 0) db->words = 1;
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
    delete db;
  }

  // everything a launch needs besides the kernel itself (kept outside the timed region)
  KParams prepare(DevBatch* db, const std::vector<uint32_t>& active, uint32_t* viol, uint32_t* err, unsigned long long* totals,
                  unsigned long long* err_totals, cudaStream_t st, size_t* smem_out) {
    const uint32_t C = prog_.nconstraints;
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
    if (active.size() != C) throw BackendError{"active mask size mismatch"};
    if (C) CK(cudaMemcpyAsync(d_active_, active.data(), (size_t)C * 4, cudaMemcpyHostToDevice, st));
    if (C) CK(cudaMemsetAsync(totals, 0, (size_t)C * sizeof(unsigned long long), st));
    if (C) CK(cudaMemsetAsync(err_totals, 0, (size_t)C * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(d_scalars_, 0, 64, st));
    size_t smem = smem_need_ + 64;
    p.smem_tables = smem <= max_smem_ ? 1u : 0u;
    if (!p.smem_tables) smem = smem_base_ + 64;
    if (smem > max_smem_) throw BackendError{"too many constraints for one launch (per-constraint counters exceed shared memory)"};
    *smem_out = smem;
    return p;
  }

  void fire(const KParams& p, uint32_t n, size_t smem, cudaStream_t st) {
    if (n == 0) return;
    // persistent-style grid: a multiple of the SM count, grid-stride over objects
    int per_sm = 8;
    if (smem > 24 * 1024) per_sm = (int)std::max<size_t>(1, (max_smem_ + 1024) / (smem + 1024));
    per_sm = std::min(per_sm, 16);
    uint32_t blocks_needed = (n + kThreads - 1) / kThreads;
    uint32_t grid = std::max(1u, std::min<uint32_t>(blocks_needed, (uint32_t)(sms_ * per_sm)));
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
    // outputs other than the kernel itself are reset outside the timed region
    size_t smem = 0;
    KParams p = prepare(db, active, db->viol, db->err, totals, err_totals, stream_, &smem);
    CK(cudaStreamSynchronize(stream_));
    // the event pair brackets exactly the evaluation kernel
    CK(cudaEventRecord(ev0_, stream_));
    fire(p, db->n, smem, stream_);
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
    fire(p, db->n, smem, st);
  }

 private:
  static constexpr uint32_t kErrCap = 1u << 20;
  void free_tables() {
    if (d_cons_) cudaFree(d_cons_);
    d_cons_ = nullptr;
    if (d_match_) cudaFree(d_match_);
    if (d_instr_) cudaFree(d_instr_);
    if (d_pool_) cudaFree(d_pool_);
    if (d_cbytes_) cudaFree(d_cbytes_);
    d_match_ = nullptr;
    d_instr_ = nullptr;
    d_pool_ = nullptr;
    d_cbytes_ = nullptr;
  }
  int device_;
  int sms_ = 148;
  size_t max_smem_ = 0, smem_need_ = 0, smem_base_ = 0;
  cudaStream_t stream_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  std::mutex mu_;
  uint64_t version_ = 0, launches_ = 0;
  GkProgram prog_{};
  GkCons* d_cons_ = nullptr;
  GkMatch* d_match_ = nullptr;
  GkInstr* d_instr_ = nullptr;
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
