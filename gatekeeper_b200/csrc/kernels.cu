// CUDA backend for sm_100a: device memory management, uploads and launches around the tile kernel
// (tile_kernel.cuh).  Host <-> device transfers are staged through pinned memory on a private stream; the
// kernel itself can also be launched on a caller-provided stream into caller-owned device buffers (multi-GPU
// gather path, where torch.distributed owns the buffers).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "backend.hpp"
#include "tile_kernel.cuh"

namespace gk {

#define CK(x)                                                                                         \
  do {                                                                                                \
    cudaError_t e_ = (x);                                                                             \
    if (e_ != cudaSuccess) throw BackendError{std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #x}; \
  } while (0)

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
  // the netlist with every slot id replaced by the slot's word offset in this batch's shared-memory slot area
  GkOp* d_ops = nullptr;
  uint32_t* d_pool = nullptr;
  GkOutEnt* d_outs = nullptr;
  uint32_t ntiles = 0, slot_words = 0, tile = 0;
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
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, gk_eval_kernel));
    max_smem_ -= fa.sharedSizeBytes;   // the kernel's few static shared bytes come out of the same per-CTA budget
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
    prog_.nphases = (uint32_t)c.phase_off.size() - 1;
    prog_.nitems = (uint32_t)c.items.size();
    if (prog_.nphases > kMaxPhases) throw BackendError{"netlist has more dependency phases than the kernel supports"};
    auto up = [&](const void* src, size_t bytes, void** dst) {
      size_t padded = (bytes + 63) / 64 * 64 + 64;
      CK(cudaMalloc(dst, padded));
      CK(cudaMemset(*dst, 0, padded));
      if (bytes) CK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    };
    up(c.ops.data(), c.ops.size() * sizeof(GkOp), (void**)&d_ops_);
    up(c.items.data(), c.items.size() * 4, (void**)&d_items_);
    up(c.outs.data(), c.outs.size() * sizeof(GkOutEnt), (void**)&d_outs_);
    up(c.phase_off.data(), c.phase_off.size() * 4, (void**)&d_phase_off_);
    up(c.match.data(), c.match.size() * sizeof(GkMatch), (void**)&d_match_);
    up(c.pool.data(), c.pool.size() * 4, (void**)&d_pool_);
    up(c.cbytes.data(), c.cbytes.size(), (void**)&d_cbytes_);
    prog_.ops = d_ops_;
    prog_.items = d_items_;
    prog_.outs = d_outs_;
    prog_.phase_off = d_phase_off_;
    prog_.match = d_match_;
    prog_.pool = d_pool_;
    prog_.cbytes = d_cbytes_;
    nscopes_ = (uint32_t)c.schema.scopes.size();
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
    PackPlan plan;
    pack_layout(hb, c, pb, plan);   // offsets only: the bytes are copied once, below, straight into pinned memory
    // ---- tiling: first row of every scope for every tile, slot offsets from the per-scope tile capacities
    const uint32_t NS = (uint32_t)c.schema.scopes.size();
    // Tile size: kTile objects.  (Shrinking tiles so that the tile count fills whole waves of resident CTAs -- 480 instead
    // of 512 objects for 1M -- measured slower: 0.816 vs 0.762 ms; the per-tile fixed cost outweighs the fuller last wave.)
    uint32_t tile = kTile;
    if (const char* ft = getenv("GK_FORCE_TILE")) tile = std::min<uint32_t>(kTile, std::max(32, atoi(ft)) / 32 * 32);
    const uint32_t ntiles = (hb.n + tile - 1) / tile;
    std::vector<uint32_t> tile_lo((size_t)(ntiles + 1) * NS), cap(NS, 0);
    for (uint32_t t = 0; t <= ntiles; ++t) {
      uint32_t* lo = &tile_lo[(size_t)t * NS];
      lo[0] = std::min<uint32_t>(t * tile, hb.n);
      for (uint32_t s = 1; s < NS; ++s) lo[s] = hb.scope_off[s][lo[c.schema.scopes[s].parent]];
      if (t)
        for (uint32_t s = 0; s < NS; ++s) cap[s] = std::max(cap[s], lo[s] - tile_lo[(size_t)(t - 1) * NS + s]);
    }
    cap[0] = tile;
    std::vector<uint32_t> slot_off(c.slot_level.size());
    uint32_t slot_words = 0;
    for (size_t i = 0; i < slot_off.size(); ++i) {
      slot_off[i] = slot_words;
      slot_words += ((cap[c.slot_level[i]] + 31) / 32 + 1 + 3) & ~3u;   // 16-byte aligned, padded: atoms store 4 words at a time
    }
    auto* db = new DevBatch();
    db->bytes = gk_align(plan.total);
    db->n = hb.n;
    db->ntiles = ntiles;
    db->tile = tile;
    db->slot_words = slot_words;
    db->prog_version = c.version;
    CK(cudaMalloc(&db->arena, db->bytes));
    CK(cudaMalloc(&db->d_tile_lo, tile_lo.size() * 4 + 64));
    // resolve slot ids -> word offsets once per batch: the kernel then addresses slots without a table lookup
    std::vector<GkOp> ops_r = c.ops;
    std::vector<uint32_t> pool_r = c.pool;
    std::vector<GkOutEnt> outs_r = c.outs;
    {
      if (slot_words > 0xffffu) throw BackendError{"slot area exceeds 16-bit word offsets"};
      std::vector<uint8_t> done(pool_r.size(), 0);
      auto so = [&](uint32_t slot) { return slot < slot_off.size() ? slot_off[slot] : 0u; };
      for (auto& op : ops_r) {
        const uint32_t kind = op.w0 & 0xffu;
        op.w0 = (op.w0 & 0xffffu) | (so(op.w0 >> 16) << 16);
        if (kind == GK_N_GATE) {
          for (uint32_t j = 0; j < op.w3; ++j) {
            uint32_t& e = pool_r[op.w1 + j];
            if (done[op.w1 + j]++) continue;
            e = (e & 0x80000000u) | so(e & 0xffffu);
          }
        } else if (kind == GK_N_BCAST || kind == GK_N_ACC) {
          for (uint32_t j = 0; j < op.w3; ++j) {
            uint32_t& e = pool_r[op.w1 + j];
            if (done[op.w1 + j]++) continue;
            e = so(e & 0xffffu) | (so(e >> 16) << 16);
          }
        } else if (kind == GK_N_MATCH) {
          op.w1 = so(op.w1 & 0xffffu);
        } else if (kind == GK_N_ATOMS) {
          for (uint32_t j = 0; j < op.w3; ++j) {
            uint32_t& e = pool_r[op.w2 + j * GK_ATOMS_ENT];
            if (done[op.w2 + j * GK_ATOMS_ENT]++) continue;
            e = (e & 0xffffu) | (so(e >> 16) << 16);
          }
        }
      }
      for (auto& oe : outs_r) {
        oe.prog_slot = (uint16_t)so(oe.prog_slot);
        oe.match_slot = (uint16_t)so(oe.match_slot);
        oe.err_slot = (uint16_t)so(oe.err_slot);
      }
    }
    auto r64 = [](size_t b) { return (b + 63) / 64 * 64 + 64; };
    CK(cudaMalloc(&db->d_ops, r64(ops_r.size() * sizeof(GkOp))));
    CK(cudaMalloc(&db->d_pool, r64(pool_r.size() * 4)));
    CK(cudaMalloc(&db->d_outs, r64(outs_r.size() * sizeof(GkOutEnt))));
    CK(cudaMemset(db->d_ops, 0, r64(ops_r.size() * sizeof(GkOp))));
    CK(cudaMemset(db->d_pool, 0, r64(pool_r.size() * 4)));
    CK(cudaMemset(db->d_outs, 0, r64(outs_r.size() * sizeof(GkOutEnt))));
    if (!ops_r.empty()) CK(cudaMemcpy(db->d_ops, ops_r.data(), ops_r.size() * sizeof(GkOp), cudaMemcpyHostToDevice));
    if (!pool_r.empty()) CK(cudaMemcpy(db->d_pool, pool_r.data(), pool_r.size() * 4, cudaMemcpyHostToDevice));
    if (!outs_r.empty()) CK(cudaMemcpy(db->d_outs, outs_r.data(), outs_r.size() * sizeof(GkOutEnt), cudaMemcpyHostToDevice));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    {
      std::lock_guard<std::mutex> l(mu_);
      // the arena image is assembled by all host threads directly in pinned memory (one copy of every byte), its
      // in-arena pointer tables are rebased to the device address, and one DMA moves it
      if (pinned_bytes_ < plan.total) {
        if (pinned_) cudaFreeHost(pinned_);
        pinned_bytes_ = plan.total + (plan.total >> 2);
        CK(cudaMallocHost(&pinned_, pinned_bytes_));
      }
      uint8_t* image = static_cast<uint8_t*>(pinned_);
      pack_copy(plan, image, host_threads_);
      db->hdr = rebase_batch(pb, image, db->arena);
      CK(cudaEventRecord(a, stream_));
      CK(cudaMemcpyAsync(db->arena, image, plan.total, cudaMemcpyHostToDevice, stream_));
      CK(cudaMemcpyAsync(db->d_tile_lo, tile_lo.data(), tile_lo.size() * 4, cudaMemcpyHostToDevice, stream_));
      CK(cudaEventRecord(b, stream_));
      CK(cudaStreamSynchronize(stream_));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (h2d_ms) *h2d_ms = ms;
    if (h2d_bytes) *h2d_bytes = plan.total + tile_lo.size() * 4 + ops_r.size() * sizeof(GkOp) + pool_r.size() * 4 + outs_r.size() * sizeof(GkOutEnt);
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
    cudaFree(db->d_ops);
    cudaFree(db->d_pool);
    cudaFree(db->d_outs);
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
    p.prog.ops = db->d_ops;      // slot ids resolved for this batch
    p.prog.pool = db->d_pool;
    p.prog.outs = db->d_outs;
    p.tile_lo = db->d_tile_lo;
    p.ntiles = db->ntiles;
    p.tile = db->tile;
    p.slot_words = db->slot_words;
    p.npeers = 0;
    p.tot_stride = 0;
    p.done_ctr = reinterpret_cast<uint32_t*>(d_scalars_) + 4;   // zeroed with the error counter before every launch
    for (int q = 0; q < GK_MAX_PEERS; ++q) {
      p.peer_viol[q] = nullptr;
      p.peer_tot[q] = nullptr;
    }
    p.timing = nullptr;
#ifdef GK_PHASE_TIMING
    if (!d_timing_) CK(cudaMalloc(&d_timing_, (kMaxPhases + 2 + 16) * 16));
    CK(cudaMemsetAsync(d_timing_, 0, (kMaxPhases + 2 + 16) * 16, st));
    p.timing = d_timing_;
#endif
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
    const size_t NS = nscopes_;
    size_t smem = 3 * r16((size_t)C * 4) + 2 * r16(NS * 4) + r16((size_t)(prog_.nphases + 1) * 4) + r16((size_t)db->slot_words * 4) + 64;
#if GK_TABLES_IN_SMEM
    smem += r16((size_t)C * sizeof(GkOutEnt)) + r16((size_t)prog_.nops * sizeof(GkOp)) + r16((size_t)prog_.nitems * 4) +
            r16((size_t)prog_.nmatch * sizeof(GkMatch)) + r16((size_t)ncols_ * sizeof(GkColumn)) + r16(NS * sizeof(GkScope)) +
            r16((size_t)prog_.npool * 4) + r16((size_t)prog_.ncbytes);
#endif
    if (smem > max_smem_)
      throw BackendError{"constraint set needs " + std::to_string(smem) + " bytes of shared memory per CTA (limit " + std::to_string(max_smem_) +
                         "): too many live netlist columns for one launch"};
    *smem_out = smem;
    return p;
  }

  void fire(const KParams& p, size_t smem, cudaStream_t st) {
    if (p.ntiles == 0) return;
    // persistent CTAs: as many as fit per SM (shared-memory bound), each walks tiles with a grid stride
    int per_sm = 1;   // resident CTAs per SM for this launch configuration (registers and shared memory)
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gk_eval_kernel, kThreads, smem));
    per_sm = std::max(1, per_sm);
    if (getenv("GK_TRACE_LAUNCH")) fprintf(stderr, "[launch] %d CTAs/SM x %d threads, %zu B smem/CTA, %u tiles of %u objects\n", per_sm, kThreads, smem, p.ntiles, p.tile);
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
#ifdef GK_PHASE_TIMING
    {
      std::vector<unsigned long long> t((kMaxPhases + 2 + 16) * 2);
      CK(cudaMemcpy(t.data(), d_timing_, t.size() * 8, cudaMemcpyDeviceToHost));
      unsigned long long tot = 0;
      for (uint32_t ph = 0; ph <= prog_.nphases; ++ph) tot += t[2 * ph];
      fprintf(stderr, "[phase timing] kernel %.3f ms, %u tiles\n", out.kernel_ms, db->ntiles);
      for (uint32_t ph = 0; ph <= prog_.nphases; ++ph)
        fprintf(stderr, "  %s %2u: %5.1f%% of CTA time, %7.0f cycles/tile, warp utilisation %4.1f%%\n", ph == prog_.nphases ? "gather" : "phase ", ph,
                100.0 * t[2 * ph] / std::max(1ull, tot), (double)t[2 * ph] / std::max(1u, db->ntiles),
                100.0 * t[2 * ph + 1] / std::max(1.0, (double)t[2 * ph] * kWarps));
      const char* kn[] = {"", "", "atom", "gate", "const", "bcast", "acc", "match", "atoms", "atoms(head)"};
      for (uint32_t k = 2; k < 10; ++k) {
        const unsigned long long cyc = t[2 * (kMaxPhases + 2) + 2 * k], cnt = t[2 * (kMaxPhases + 2) + 2 * k + 1];
        if (cnt) fprintf(stderr, "  items %-11s: %6.1f per tile, %7.0f cycles each, %8.0f warp-cycles per tile\n", kn[k], (double)cnt / db->ntiles,
                         (double)cyc / cnt, (double)cyc / db->ntiles);
      }
    }
#endif
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

  void* ingest(const IngestReq&, IngestStats*, std::vector<uint32_t>*) override { throw BackendError{"device ingest: not built yet"}; }

  void eval_into(void* b, const std::vector<uint32_t>& active, const DevOutPtrs& dst) override {
    auto* db = static_cast<DevBatch*>(b);
    std::lock_guard<std::mutex> l(mu_);
    CK(cudaSetDevice(device_));
    size_t smem = 0;
    cudaStream_t st = static_cast<cudaStream_t>(dst.stream);
    KParams p = prepare(db, active, static_cast<uint32_t*>(dst.viol), static_cast<uint32_t*>(dst.err), static_cast<unsigned long long*>(dst.totals),
                        static_cast<unsigned long long*>(dst.err_totals), st, &smem);
    if (dst.npeers > GK_MAX_PEERS) throw BackendError{"too many peers for the fused exchange"};
    p.npeers = dst.npeers;
    p.tot_stride = dst.tot_stride;
    for (uint32_t q = 0; q < dst.npeers; ++q) {
      p.peer_viol[q] = reinterpret_cast<uint32_t*>(dst.peer_viol[q]);
      p.peer_tot[q] = reinterpret_cast<unsigned long long*>(dst.peer_tot[q]);
    }
    fire(p, smem, st);
  }

 private:
  static constexpr uint32_t kErrCap = 1u << 20;
  void free_tables() {
    void** ptrs[] = {(void**)&d_outs_, (void**)&d_ops_, (void**)&d_items_, (void**)&d_phase_off_, (void**)&d_match_, (void**)&d_pool_, (void**)&d_cbytes_};
    for (auto pp : ptrs) {
      if (*pp) cudaFree(*pp);
      *pp = nullptr;
    }
  }
  int device_;
  int sms_ = 148;
  size_t max_smem_ = 0, sm_smem_ = 0;
  cudaStream_t stream_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  std::mutex mu_;
  uint64_t version_ = 0, launches_ = 0;
  GkProgram prog_{};
  std::vector<uint32_t> last_active_;
  uint32_t ncols_ = 0, nscopes_ = 0;
  GkOp* d_ops_ = nullptr;
  uint32_t* d_items_ = nullptr;
  GkOutEnt* d_outs_ = nullptr;
  uint32_t* d_phase_off_ = nullptr;
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
  unsigned long long* d_timing_ = nullptr;
  int host_threads_ = effective_cpus();
  void* pinned_ = nullptr;
  size_t pinned_bytes_ = 0;
};

Backend* make_backend(int device) { return new CudaBackend(device); }

}  // namespace gk
