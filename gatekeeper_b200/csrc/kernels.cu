// CUDA backend for sm_100a: device memory management, uploads and launches around the tile kernel
// (tile_kernel.cuh).  Host <-> device transfers are staged through pinned memory on a private stream; the
// kernel itself can also be launched on a caller-provided stream into caller-owned device buffers (multi-GPU
// gather path, where torch.distributed owns the buffers).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <functional>
#include <memory>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <dlfcn.h>

#include <list>
#include <sstream>

#include "backend.hpp"
#include "ingest_kernels.cuh"
#include "spec_codegen.hpp"
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
  uint32_t* d_tile_list = nullptr;   // [ntiles] tiles the specialised kernel hands to the interpreter (objects too big for its mask registers)
  // the netlist with every slot id replaced by the slot's word offset in this batch's shared-memory slot area
  GkOp* d_ops = nullptr;
  uint32_t* d_pool = nullptr;
  GkOutEnt* d_outs = nullptr;
  uint32_t ntiles = 0, slot_words = 0, tile = 0;
  uint64_t prog_version = 0;
  bool gvk_uniform = false;   // every object that was not skipped has the same apiVersion and kind
  std::vector<uint32_t> cap;  // rows of the largest tile per scope (sizes the slot area of any netlist over this batch)
  bool fork = false;          // a view made by fork_batch: arena and tiling belong to the batch it was forked from
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
    CK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    {
      // The tokeniser of the NEXT page runs on a high-priority stream: its launches are small (405 CTAs per 32 MB chunk, 15 % of
      // the warp slots) and would otherwise only start in the tails of the current page's extraction kernels, whose grids
      // keep the block scheduler busy -- measured as fully serialised (12.5 + 11.9 = 24.4 ms per page).
      int lo = 0, hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      const bool prio = !getenv("GK_NO_STREAM_PRIORITY");
      CK(cudaStreamCreateWithPriority(&front_stream_, cudaStreamNonBlocking, prio ? hi : lo));
    }
    // result bitmaps are copied back into page-locked blocks (backend.hpp HostBlockAlloc)
    host_block_hooks().alloc = [](size_t bytes) -> void* {
      void* p = nullptr;
      return cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess ? p : nullptr;
    };
    host_block_hooks().release = [](void* p) { cudaFreeHost(p); };
    {
      // per-batch device memory comes from the stream-ordered pool and goes back to it: with the release threshold lifted a
      // resident-batch-sized arena is reused by the next batch instead of being mapped / unmapped by the driver every time
      cudaMemPool_t pool;
      CK(cudaDeviceGetDefaultMemPool(&pool, device_));
      unsigned long long keep = ~0ull;
      CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
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
  const char* last_kernel() const override { return last_kernel_; }

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
    spec_select(c);
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
    uint32_t ntiles = 0;
    std::vector<uint32_t> tile_lo, cap;
    for (;; tile = std::max(32u, tile / 2 / 32 * 32)) {
      ntiles = (hb.n + tile - 1) / tile;
      tile_lo.assign((size_t)(ntiles + 1) * NS, 0);
      cap.assign(NS, 0);
      for (uint32_t t = 0; t <= ntiles; ++t) {
        uint32_t* lo = &tile_lo[(size_t)t * NS];
        lo[0] = std::min<uint32_t>(t * tile, hb.n);
        for (uint32_t s = 1; s < NS; ++s) lo[s] = hb.scope_off[s][lo[c.schema.scopes[s].parent]];
        if (t)
          for (uint32_t s = 0; s < NS; ++s) cap[s] = std::max(cap[s], lo[s] - tile_lo[(size_t)(t - 1) * NS + s]);
      }
      if (tile <= 32 || tile_fits(c, cap, tile)) break;   // (a 32-object tile that still does not fit is refused at launch)
    }
    auto* db = new DevBatch();
    db->bytes = gk_align(plan.total);
    db->n = hb.n;
    db->ntiles = ntiles;
    db->tile = tile;
    db->prog_version = c.version;
    db->gvk_uniform = hb.gvk_hi == 0 || hb.gvk_lo == hb.gvk_hi;
    dmalloc(&db->arena, db->bytes);
    dmalloc(&db->d_tile_lo, tile_lo.size() * 4 + 64);
    finish_batch(db, c, cap);
    const size_t tables_bytes = c.ops.size() * sizeof(GkOp) + c.pool.size() * 4 + c.outs.size() * sizeof(GkOutEnt);
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
    if (h2d_bytes) *h2d_bytes = plan.total + tile_lo.size() * 4 + tables_bytes;
    return db;
  }

  // ---- adaptive tile size: every bit column of a tile lives in the CTA's shared memory, so a page whose objects iterate very
  // many rows (one Pod with thousands of containers) needs smaller tiles, not a refusal
  static uint32_t slot_words_of(const Compiled& c, std::vector<uint32_t> cap, uint32_t tile) {
    cap[0] = tile;
    uint64_t w = 0;
    for (size_t i = 0; i < c.slot_level.size(); ++i) w += ((cap[c.slot_level[i]] + 31) / 32 + 1 + 3) & ~3u;
    return (uint32_t)std::min<uint64_t>(w, 0xffffffffu);
  }
  size_t smem_of(const Compiled& c, uint32_t slot_words) const {
    auto r16 = [](size_t x) { return (x + 15) / 16 * 16; };
    const size_t C = c.cons_match.size(), NS = c.schema.scopes.size();
    size_t smem = 3 * r16(C * 4) + 2 * r16(NS * 4) + r16(c.phase_off.size() * 4) + r16((size_t)slot_words * 4) + 64;
#if GK_TABLES_IN_SMEM
    smem += r16(C * sizeof(GkOutEnt)) + r16(c.ops.size() * sizeof(GkOp)) + r16(c.items.size() * 4) + r16(c.match.size() * sizeof(GkMatch)) +
            r16(c.schema.cols.size() * sizeof(GkColumn)) + r16(NS * sizeof(GkScope)) + r16(c.pool.size() * 4) + r16(c.cbytes.size());
#endif
    return smem;
  }
  bool tile_fits(const Compiled& c, const std::vector<uint32_t>& cap, uint32_t tile) const {
    const uint32_t sw = slot_words_of(c, cap, tile);
    if (sw > 0xffffu || smem_of(c, sw) > max_smem_) return false;
    return !c.amb || (slot_words_of(*c.amb, cap, tile) <= 0xffffu && smem_of(*c.amb, slot_words_of(*c.amb, cap, tile)) <= max_smem_);
  }

  // slot area of a tile from the per-scope tile capacities; the netlist with slot ids resolved to word offsets; output planes
  void finish_batch(DevBatch* db, const Compiled& c, std::vector<uint32_t> cap) {
    cap[0] = db->tile;
    db->cap = cap;
    std::vector<uint32_t> slot_off(c.slot_level.size());
    uint32_t slot_words = 0;
    for (size_t i = 0; i < slot_off.size(); ++i) {
      slot_off[i] = slot_words;
      slot_words += ((cap[c.slot_level[i]] + 31) / 32 + 1 + 3) & ~3u;   // 16-byte aligned, padded: atoms store 4 words at a time
    }
    db->slot_words = slot_words;
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
        } else if (kind == GK_N_BCAST || kind == GK_N_ACC || kind == GK_N_ACC2) {
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
    dmalloc(&db->d_ops, r64(ops_r.size() * sizeof(GkOp)));
    dmalloc(&db->d_pool, r64(pool_r.size() * 4));
    dmalloc(&db->d_outs, r64(outs_r.size() * sizeof(GkOutEnt)));
    CK(cudaMemsetAsync(db->d_ops, 0, r64(ops_r.size() * sizeof(GkOp)), stream_));
    CK(cudaMemsetAsync(db->d_pool, 0, r64(pool_r.size() * 4), stream_));
    CK(cudaMemsetAsync(db->d_outs, 0, r64(outs_r.size() * sizeof(GkOutEnt)), stream_));
    push_small(db->d_ops, ops_r.data(), ops_r.size() * sizeof(GkOp));
    push_small(db->d_pool, pool_r.data(), pool_r.size() * 4);
    push_small(db->d_outs, outs_r.data(), outs_r.size() * sizeof(GkOutEnt));
    db->words = std::max<uint32_t>(1, (uint32_t)((c.cons_match.size() + 31) / 32));
    dmalloc(&db->d_tile_list, (size_t)(db->ntiles + 1) * 4);
    dmalloc(&db->viol, (size_t)std::max(db->n, 1u) * db->words * 4);
    dmalloc(&db->err, (size_t)std::max(db->n, 1u) * db->words * 4);
    CK(cudaStreamSynchronize(stream_));
  }

  // ---- small uploads through the SMs (gk_push_kernel): staged in a ring of mapped page-locked memory.  The caller's buffer is free
  // to go as soon as the call returns.
  void push_small(void* dst, const void* src, size_t bytes) {
    if (!bytes) return;
    std::lock_guard<std::mutex> sl(stage_mu_);   // (uploads and forks of different batches run concurrently: one ring, one writer)
    if (bytes > (4u << 20) || getenv("GK_NO_PUSH_KERNEL")) {   // (big: the copy engine is the right tool)
      CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream_));
      CK(cudaStreamSynchronize(stream_));
      return;
    }
    const size_t need = (bytes + 15) & ~(size_t)15;
    if (!stage_host_) {
      stage_cap_ = 16u << 20;
      CK(cudaHostAlloc(reinterpret_cast<void**>(&stage_host_), stage_cap_, cudaHostAllocMapped | cudaHostAllocPortable));
      CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&stage_dev_), stage_host_, 0));
      stage_off_ = 0;
    }
    if (stage_off_ + need > stage_cap_) {   // wrap: everything staged so far must have been read
      CK(cudaStreamSynchronize(stream_));
      stage_off_ = 0;
    }
    memcpy(stage_host_ + stage_off_, src, bytes);
    const uint32_t grid = (uint32_t)std::min<size_t>(64, (bytes + 4095) / 4096);
    gk_push_kernel<<<grid, 256, 0, stream_>>>(static_cast<uint8_t*>(dst), stage_dev_ + stage_off_, bytes);
    CK(cudaGetLastError());
    stage_off_ += need;
  }
  uint8_t *stage_host_ = nullptr, *stage_dev_ = nullptr;
  size_t stage_cap_ = 0, stage_off_ = 0;
  std::mutex stage_mu_;

  template <class T>
  void dmalloc(T** p, size_t bytes) { CK(cudaMallocAsync(reinterpret_cast<void**>(p), std::max<size_t>(bytes, 256), stream_)); }
  void dfree(void* p) {
    if (p) cudaFreeAsync(p, stream_);
  }
  void release(void* b) override {
    auto* db = static_cast<DevBatch*>(b);
    if (!db) return;
    cudaSetDevice(device_);
    if (!db->fork) dfree(db->arena);
    dfree(db->viol);
    dfree(db->err);
    if (!db->fork) dfree(db->d_tile_lo);
    dfree(db->d_tile_list);
    dfree(db->d_ops);
    dfree(db->d_pool);
    dfree(db->d_outs);
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
    p.tile_list = db->d_tile_list;
    p.tile_count = reinterpret_cast<uint32_t*>(d_scalars_) + 8;   // zeroed with the error counter before every launch
    p.list_mode = 0;
    p.pad_ = 0;
    spec_next_ = spec_ready(db->n);   // (an NVRTC run the first time: here, outside the timed region)
#ifdef GK_PHASE_TIMING
    if (!d_timing_) CK(cudaMalloc(&d_timing_, (kMaxPhases + 2 + 16) * 16));
    CK(cudaMemsetAsync(d_timing_, 0, (kMaxPhases + 2 + 16) * 16, st));
    p.timing = d_timing_;
#endif
    if (active.size() != C) throw BackendError{"active mask size mismatch"};
    if (db->words == 2 && ((reinterpret_cast<uintptr_t>(viol) | reinterpret_cast<uintptr_t>(err)) & 7u))
      throw BackendError{"result bitmaps must be 8-byte aligned (two-word rows are stored as one 64-bit word)"};
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

  void fire(KParams p, size_t smem, cudaStream_t st) {
    if (p.ntiles == 0) return;
    // persistent CTAs: as many as fit per SM (shared-memory bound), each walks tiles with a grid stride
    int per_sm = 1;   // resident CTAs per SM for this launch configuration (registers and shared memory)
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gk_eval_kernel, kThreads, smem));
    per_sm = std::max(1, per_sm);
    if (getenv("GK_TRACE_LAUNCH")) fprintf(stderr, "[launch] %d CTAs/SM x %d threads, %zu B smem/CTA, %u tiles of %u objects\n", per_sm, kThreads, smem, p.ntiles, p.tile);
    uint32_t grid = std::max(1u, std::min<uint32_t>(p.ntiles, (uint32_t)(sms_ * per_sm)));
    last_kernel_ = "gk_eval_kernel";
    if (spec_next_) {
      // the kernel generated for this constraint set does the tiles whose objects fit its mask registers (all of them, normally);
      // the interpreter behind it takes the tiles it listed and, in the fused exchange, publishes the totals of both
      SpecEntry& se = *spec_cur_;
      void* args[] = {&p};
      CK(cudaLaunchKernel(reinterpret_cast<const void*>(se.kern), dim3(p.ntiles), dim3((unsigned)se.threads), args, se.src.smem, st));
      ++launches_;
      p.list_mode = 1;
      last_kernel_ = "gk_spec_kernel";
    }
    gk_eval_kernel<<<grid, kThreads, smem, st>>>(p);
    CK(cudaGetLastError());
    ++launches_;
  }

  // ---- the specialised kernel: source per constraint-set version (spec_codegen.cpp), compiled by NVRTC when first needed
  struct SpecEntry {
    uint64_t version = 0;
    SpecSource src;
    int state = 0;            // 0 = not compiled yet, 1 = ready, -1 = failed (the interpreter runs; reported once on stderr)
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kern = nullptr;
    int threads = 128;
    double compile_ms = 0;
  };
  void spec_select(const Compiled& c) {
    spec_cur_ = nullptr;
    if (!spec_on_) return;
    for (auto it = spec_cache_.begin(); it != spec_cache_.end(); ++it)
      if (it->version == c.version) {
        spec_cache_.splice(spec_cache_.begin(), spec_cache_, it);
        spec_cur_ = &spec_cache_.front();
        return;
      }
    spec_cache_.emplace_front();
    spec_cache_.front().version = c.version;
    spec_cache_.front().src = spec_codegen(c);
    spec_cur_ = &spec_cache_.front();
    while (spec_cache_.size() > 8) {
      if (spec_cache_.back().lib) cudaLibraryUnload(spec_cache_.back().lib);
      spec_cache_.pop_back();
    }
  }
  bool spec_ready(uint32_t n) {
    if (!spec_cur_ || n < spec_min_objects_) return false;
    if (spec_cur_->state == 0) spec_compile(*spec_cur_);
    return spec_cur_->state == 1;
  }
  void spec_compile(SpecEntry& se) {
    se.state = -1;
    const auto t0 = std::chrono::steady_clock::now();
    try {
      // NVRTC is loaded at run time: the library itself links only the CUDA runtime (and loads in a container without a driver)
      typedef void* Prog;
      typedef int (*CreateFn)(Prog*, const char*, const char*, int, const char* const*, const char* const*);
      typedef int (*CompileFn)(Prog, int, const char* const*);
      typedef int (*SizeFn)(Prog, size_t*);
      typedef int (*GetFn)(Prog, char*);
      typedef int (*DestroyFn)(Prog*);
      static void* h = nullptr;
      if (!h) {
        for (const char* name : {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"}) {
          h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
          if (h) break;
        }
      }
      if (!h) throw BackendError{"libnvrtc.so.12 not found"};
      auto create = reinterpret_cast<CreateFn>(dlsym(h, "nvrtcCreateProgram"));
      auto compile = reinterpret_cast<CompileFn>(dlsym(h, "nvrtcCompileProgram"));
      auto log_size = reinterpret_cast<SizeFn>(dlsym(h, "nvrtcGetProgramLogSize"));
      auto get_log = reinterpret_cast<GetFn>(dlsym(h, "nvrtcGetProgramLog"));
      auto bin_size = reinterpret_cast<SizeFn>(dlsym(h, "nvrtcGetCUBINSize"));
      auto get_bin = reinterpret_cast<GetFn>(dlsym(h, "nvrtcGetCUBIN"));
      auto destroy = reinterpret_cast<DestroyFn>(dlsym(h, "nvrtcDestroyProgram"));
      if (!create || !compile || !log_size || !get_log || !bin_size || !get_bin || !destroy) throw BackendError{"NVRTC entry points missing"};
      Prog prog = nullptr;
      if (create(&prog, se.src.src.c_str(), "gk_spec_kernel.cu", 0, nullptr, nullptr) != 0) throw BackendError{"nvrtcCreateProgram failed"};
      const std::string threads = "-DGK_SPEC_THREADS=" + std::to_string(spec_threads_), minb = "-DGK_SPEC_MINB=" + std::to_string(spec_minb_);
      std::vector<std::string> extra;   // GK_SPEC_DEFS="-DX -DY": measurement switches of the generated text
      if (const char* d = getenv("GK_SPEC_DEFS")) {
        std::istringstream in(d);
        std::string tok;
        while (in >> tok) extra.push_back(tok);
      }
      std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device", threads.c_str(), minb.c_str()};
      for (auto& x : extra) opts.push_back(x.c_str());
      const int rc = compile(prog, (int)opts.size(), opts.data());
      if (rc != 0) {
        size_t ls = 0;
        log_size(prog, &ls);
        std::string log(ls + 1, 0);
        if (ls) get_log(prog, &log[0]);
        destroy(&prog);
        throw BackendError{"NVRTC: " + log.substr(0, 2000)};
      }
      size_t bs = 0;
      bin_size(prog, &bs);
      std::vector<char> cubin(bs);
      get_bin(prog, cubin.data());
      destroy(&prog);
      CK(cudaLibraryLoadData(&se.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
      CK(cudaLibraryGetKernel(&se.kern, se.lib, "gk_spec_kernel"));
      se.threads = spec_threads_;
      if (se.src.smem > 48 * 1024)
        CK(cudaFuncSetAttribute(reinterpret_cast<const void*>(se.kern), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)se.src.smem));
      se.state = 1;
    } catch (const BackendError& e) {
      fprintf(stderr, "[gatekeeper_b200] the kernel generated for constraint set %llu could not be built (%s): the netlist interpreter runs instead\n",
              (unsigned long long)se.version, e.msg.c_str());
    }
    se.compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("GK_TRACE_LAUNCH"))
      fprintf(stderr, "[spec] constraint set %llu: %zu bytes of source, %zu atoms as immediates / %zu generic, NVRTC %.0f ms, state %d\n", (unsigned long long)se.version,
              se.src.src.size(), se.src.n_fast, se.src.n_generic, se.compile_ms, se.state);
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
        CK(cudaMemcpyAsync(out.viol.data(), db->viol, out.viol.size() * 4, cudaMemcpyDeviceToHost, stream_));
        CK(cudaMemcpyAsync(out.err.data(), db->err, out.err.size() * 4, cudaMemcpyDeviceToHost, stream_));
        CK(cudaStreamSynchronize(stream_));
      }
    }
  }

  // ---- device ingest (ingest_core.h / ingest_kernels.cuh): raw JSON blob -> resident columnar batch
  struct Scratch {   // grow-only device buffer
    uint8_t* p = nullptr;
    size_t cap = 0;
    uint8_t* need(size_t bytes) {
      if (bytes > cap) {
        if (p) cudaFree(p);
        p = nullptr;
        cap = bytes + (bytes >> 3) + 4096;
        cudaError_t e = cudaMalloc(&p, cap);
        if (e != cudaSuccess) {
          cap = 0;
          throw BackendError{std::string("CUDA error: ") + cudaGetErrorString(e) + " allocating ingest scratch"};
        }
      }
      return p;
    }
  };
  struct Carver {   // 256-byte aligned sub-allocations of one buffer
    size_t off = 0;
    size_t take(size_t bytes) {
      size_t o = gk_align(off);
      off = o + std::max<size_t>(bytes, 16);
      return o;
    }
  };

  void upload_htab(const HashTabHost& h, unsigned long long** dk, uint32_t** dv, size_t* dcap) {
    const size_t cap = (size_t)h.mask + 1;
    if (*dcap != cap) {
      if (*dk) cudaFree(*dk);
      if (*dv) cudaFree(*dv);
      CK(cudaMalloc(dk, cap * 8));
      CK(cudaMalloc(dv, cap * 4));
      *dcap = cap;
    }
    CK(cudaMemcpyAsync(*dk, h.keys.data(), cap * 8, cudaMemcpyHostToDevice, stream_));
    CK(cudaMemcpyAsync(*dv, h.vals.data(), cap * 4, cudaMemcpyHostToDevice, stream_));
    CK(cudaStreamSynchronize(stream_));
  }

  void* ingest(const IngestReq& rq, IngestStats* st, std::vector<uint32_t>* status) override {
    std::lock_guard<std::mutex> l(ingest_mu_);
    CK(cudaSetDevice(device_));
    const auto T0 = std::chrono::steady_clock::now();
    const XProgHost& xh = *rq.xprog;
    const Compiled& c = *rq.c;
    const uint32_t n = (uint32_t)rq.n, NS = (uint32_t)xh.scopes.size(), NK = xh.ncounters(), NC = (uint32_t)xh.cols.size();
    const unsigned long long B = rq.ooff[n];
    if (rq.n > 0x7fffff00ull) throw BackendError{"device ingest: more than 2^31 objects in one batch"};
    // ---- lookup tables
    if (sid_.nstrings != rq.strings->size()) {
      build_sid_table(*rq.strings, sid_);
      upload_htab(sid_.tab, &d_sid_keys_, &d_sid_vals_, &d_sid_cap_);
    }
    if (lut_.mask == 0) {
      lut_.init(1u << 20);
      upload_htab(lut_, &d_lut_keys_, &d_lut_vals_, &d_lut_cap_);
    }
    // ---- extraction program + namespace cache tables: one host image, one copy
    std::vector<uint32_t> xkeys = xh.xkeys;
    std::vector<uint8_t> xbytes = xh.xbytes;
    const uint32_t excl_off = (uint32_t)xkeys.size();
    uint32_t excl_n = 0;
    for (auto& pat : rq.excluded) {
      const bool pre = !pat.empty() && pat.front() == '*', suf = pat.size() > (pre ? 1u : 0u) && pat.back() == '*';
      const std::string core = pat.substr(pre ? 1 : 0, pat.size() - (pre ? 1 : 0) - (suf ? 1 : 0));
      xkeys.push_back(pre && suf ? GK_W_CONTAINS : pre ? GK_W_SUFFIX : suf ? GK_W_PREFIX : GK_W_EXACT);
      xkeys.push_back((uint32_t)xbytes.size());
      xkeys.push_back((uint32_t)core.size());
      xbytes.insert(xbytes.end(), core.begin(), core.end());
      ++excl_n;
    }
    const NsTableHost& ns = *rq.ns;
    Carver tc;
    const size_t o_cl = tc.take(xh.cl.size() * sizeof(GkXClosure)), o_cols = tc.take(xh.cols.size() * sizeof(GkXCol)),
                 o_scopes = tc.take(xh.scopes.size() * sizeof(GkXScope)), o_order = tc.take(xh.col_order.size() * 4), o_xkeys = tc.take(xkeys.size() * 4),
                 o_xargs = tc.take(xh.xargs.size() * 4), o_xbytes = tc.take(xbytes.size()), o_nskeys = tc.take(ns.tab.keys.size() * 8),
                 o_nsvals = tc.take(ns.tab.vals.size() * 4), o_nsnoff = tc.take(ns.nsn_off.size() * 4), o_nsnbytes = tc.take(ns.nsn_bytes.size());
    const size_t tab_bytes = gk_align(tc.off);
    std::vector<uint8_t> timg(tab_bytes, 0);
    auto putv = [&](size_t off, const void* src, size_t bytes) {
      if (bytes) memcpy(timg.data() + off, src, bytes);
    };
    putv(o_cl, xh.cl.data(), xh.cl.size() * sizeof(GkXClosure));
    putv(o_cols, xh.cols.data(), xh.cols.size() * sizeof(GkXCol));
    putv(o_scopes, xh.scopes.data(), xh.scopes.size() * sizeof(GkXScope));
    putv(o_order, xh.col_order.data(), xh.col_order.size() * 4);
    putv(o_xkeys, xkeys.data(), xkeys.size() * 4);
    putv(o_xargs, xh.xargs.data(), xh.xargs.size() * 4);
    putv(o_xbytes, xbytes.data(), xbytes.size());
    putv(o_nskeys, ns.tab.keys.data(), ns.tab.keys.size() * 8);
    putv(o_nsvals, ns.tab.vals.data(), ns.tab.vals.size() * 4);
    putv(o_nsnoff, ns.nsn_off.data(), ns.nsn_off.size() * 4);
    putv(o_nsnbytes, ns.nsn_bytes.data(), ns.nsn_bytes.size());
    uint8_t* d_tab = tabs_.need(tab_bytes);
    push_small(d_tab, timg.data(), tab_bytes);
    GkXProg xp;
    memset(&xp, 0, sizeof xp);
    xp.cl = reinterpret_cast<const GkXClosure*>(d_tab + o_cl);
    xp.cols = reinterpret_cast<const GkXCol*>(d_tab + o_cols);
    xp.scopes = reinterpret_cast<const GkXScope*>(d_tab + o_scopes);
    xp.col_order = reinterpret_cast<const uint32_t*>(d_tab + o_order);
    xp.xkeys = reinterpret_cast<const uint32_t*>(d_tab + o_xkeys);
    xp.xargs = reinterpret_cast<const uint32_t*>(d_tab + o_xargs);
    xp.xbytes = d_tab + o_xbytes;
    xp.ncl = (uint32_t)xh.cl.size();
    xp.ncols = NC;
    xp.nscopes = NS;
    xp.nbytecols = xh.nbytecols;
    xp.sid_tab.keys = d_sid_keys_;
    xp.sid_tab.vals = d_sid_vals_;
    xp.sid_tab.mask = sid_.tab.mask;
    xp.sid_true = sid_.sid_true;
    xp.sid_false = sid_.sid_false;
    xp.sid_null = sid_.sid_null;
    xp.ns_tab.keys = reinterpret_cast<unsigned long long*>(d_tab + o_nskeys);
    xp.ns_tab.vals = reinterpret_cast<uint32_t*>(d_tab + o_nsvals);
    xp.ns_tab.mask = ns.tab.mask;
    xp.nsn_off = reinterpret_cast<const uint32_t*>(d_tab + o_nsnoff);
    xp.nsn_bytes = d_tab + o_nsnbytes;
    xp.excl_off = excl_off;
    xp.excl_n = excl_n;
    // ---- front buffer (blob, offsets, tape): the prefetched copy of this page, or copy + tokenise now
    uint32_t clanes = 1;
    if (const char* ev = getenv("GK_INGEST_COL_LANES")) clanes = std::max(1, std::min(32, atoi(ev)));
    while (clanes & (clanes - 1)) --clanes;
    cudaEvent_t e0, e1, e2, e3;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventCreate(&e2));
    CK(cudaEventCreate(&e3));
    CK(cudaEventRecord(e0, stream_));
    Front* fr = nullptr;
    {
      std::lock_guard<std::mutex> fl(front_mu_);
      for (auto& f : fronts_)
        if (f.pending && f.host_blob == rq.blob && f.n == rq.n) fr = &f;   // (prefetched)
      if (!fr) fr = &start_front(rq.blob, rq.ooff, rq.n);
      fr->pending = false;
    }
    CK(cudaStreamWaitEvent(stream_, fr->done, 0));
    // ---- back scratch: header counters + scratch flags, totals, scan block sums, miss list, the pointer tables of the count phase
    const uint32_t miss_cap = 1u << 17;
    const uint32_t NT = GK_CNT_EXTRA + NS + xh.nbytecols + 4;          // totals: header counters, scopes, byte columns
    Carver sc;
    const size_t o_counts = sc.take((size_t)GK_CNT_EXTRA * n * 4 + 16), o_flags = sc.take((size_t)n * 4 + 16), o_totals = sc.take((size_t)(NT + NS + 4) * 4),
                 o_sums = sc.take((size_t)NT * 4096 * 4), o_miss = sc.take((size_t)miss_cap * sizeof(GkMiss)), o_fill = sc.take((size_t)miss_cap * 8),
                 o_prec = sc.take((size_t)NS * 8);
    uint8_t* d_s = scratch_.need(gk_align(sc.off));
    GkIngestIn in = fr->in;
    in.source = rq.source;
    in.counts = reinterpret_cast<uint32_t*>(d_s + o_counts);
    in.misses = reinterpret_cast<GkMiss*>(d_s + o_miss);
    uint32_t* d_totals = reinterpret_cast<uint32_t*>(d_s + o_totals);   // [0,4) header, [4, 4+NS) scopes, then byte columns
    uint32_t* d_cap = d_totals + NT;
    in.nmiss = d_cap + NS;
    in.miss_cap = miss_cap;
    uint32_t* d_fill = reinterpret_cast<uint32_t*>(d_s + o_fill);
    uint32_t* d_sums = reinterpret_cast<uint32_t*>(d_s + o_sums);
    // exclusive scan of a[0..len) in place (+ a[len] = total when `tail`), total also at d_totals[slot]
    auto scan = [&](uint32_t* a, uint32_t len, uint32_t slot, bool tail) {
      const uint32_t nb = (len + kScanBlock - 1) / kScanBlock;
      if (nb > 4096) throw BackendError{"device ingest: more than 16M rows in one scope of one batch"};
      uint32_t* sums = d_sums + (size_t)slot * 4096;
      if (nb) gk_scan_sums_kernel<<<nb, kScanThreads, 0, stream_>>>(a, len, sums);
      gk_scan_top_kernel<<<1, kScanThreads, 0, stream_>>>(sums, nb, d_totals + slot);
      if (nb) gk_scan_apply_kernel<<<nb, kScanThreads, 0, stream_>>>(a, len, sums, d_totals + slot, tail ? 1u : 0u);
      else if (tail) CK(cudaMemsetAsync(a, 0, 4, stream_));
      launches_ += nb ? 3 : 1;
    };
    auto blocks = [](uint32_t rows) { return (rows + kIngestThreads - 1) / kIngestThreads; };
    // the count phase works on scratch: flags, CSR offsets (= scanned member counts), generator nodes, row handles, byte lengths
    GkIngestOut outc;
    memset(&outc, 0, sizeof outc);
    outc.flags = reinterpret_cast<uint32_t*>(d_s + o_flags);
    outc.row_rec = reinterpret_cast<GkRowRec* const*>(d_s + o_prec);
    std::vector<uint64_t> h_prec(NS, 0);
    std::vector<uint32_t*> d_cnt(NS, nullptr), d_coll(NS, nullptr), d_blen(xh.nbytecols, nullptr);
    std::vector<void*> temps;   // stream-ordered allocations of this call
    struct TempGuard {
      CudaBackend* be;
      std::vector<void*>* v;
      ~TempGuard() {
        for (void* p2 : *v) be->dfree(p2);
      }
    } temp_guard{this, &temps};
    auto talloc = [&](size_t bytes) {
      uint8_t* p2 = nullptr;
      dmalloc(&p2, bytes);
      temps.push_back(p2);
      return p2;
    };
    CK(cudaEventRecord(e1, stream_));
    std::vector<uint32_t> total(NS, 0), htot(GK_CNT_EXTRA, 0), btot(xh.nbytecols, 0);
    total[0] = n;
    if (n) {
      gk_hcount_kernel<<<blocks(n), kIngestThreads, 0, stream_>>>(xp, in, outc);
      ++launches_;
      for (uint32_t k = 0; k < GK_CNT_EXTRA; ++k) scan(in.counts + (size_t)k * n, n, k, false);
      // scopes, one depth level at a time (the rows of a level size the kernels of the next)
      uint32_t maxd = 0;
      std::vector<uint32_t> depth(NS, 0);
      for (uint32_t s2 = 1; s2 < NS; ++s2) depth[s2] = depth[xh.scopes[s2].parent] + 1, maxd = std::max(maxd, depth[s2]);
      std::vector<uint32_t> hbuf(NT, 0);
      for (uint32_t d = 1; d <= maxd; ++d) {
        for (uint32_t t = 1; t < NS; ++t) {
          if (depth[t] != d) continue;
          const uint32_t prows = total[xh.scopes[t].parent];
          d_cnt[t] = reinterpret_cast<uint32_t*>(talloc(((size_t)prows + 1) * 4));
          d_coll[t] = reinterpret_cast<uint32_t*>(talloc(((size_t)prows + 1) * 4));
          if (prows) {
            gk_scope_count_kernel<<<blocks(prows), kIngestThreads, 0, stream_>>>(xp, in, outc, t, prows, d_cnt[t], d_coll[t]);
            ++launches_;
          }
          scan(d_cnt[t], prows, GK_CNT_EXTRA + t, true);
        }
        CK(cudaMemcpyAsync(hbuf.data(), d_totals, (size_t)NT * 4, cudaMemcpyDeviceToHost, stream_));
        CK(cudaStreamSynchronize(stream_));
        CK(cudaGetLastError());
        bool any = false;
        for (uint32_t t = 1; t < NS; ++t) {
          if (depth[t] != d) continue;
          total[t] = hbuf[GK_CNT_EXTRA + t];
          h_prec[t] = reinterpret_cast<uint64_t>(talloc(((size_t)total[t] + 1) * sizeof(GkRowRec)));
          any = true;
        }
        if (!any) continue;
        push_small(d_s + o_prec, h_prec.data(), (size_t)NS * 8);
        for (uint32_t t = 1; t < NS; ++t) {
          if (depth[t] != d) continue;
          const uint32_t prows = total[xh.scopes[t].parent];
          if (!prows || !total[t]) continue;
          gk_scope_fill_kernel<<<blocks(prows), kIngestThreads, 0, stream_>>>(xp, in, outc, t, prows, d_cnt[t], d_coll[t]);
          ++launches_;
        }
      }
      // byte-encoded columns: decoded length per row -> offsets
      for (uint32_t ci = 0; ci < NC; ++ci) {
        const GkXCol& xc = xh.cols[ci];
        if (!(xc.enc & GK_ENC_BYTES)) continue;
        const uint32_t rows = total[xc.scope];
        d_blen[xc.bytes_slot] = reinterpret_cast<uint32_t*>(talloc(((size_t)rows + 1) * 4));
        if (rows) {
          gk_bcol_len_kernel<<<blocks(rows), kIngestThreads, 0, stream_>>>(xp, in, outc, ci, rows, d_blen[xc.bytes_slot]);
          ++launches_;
        }
        scan(d_blen[xc.bytes_slot], rows, GK_CNT_EXTRA + NS + xc.bytes_slot, true);
      }
      CK(cudaMemcpyAsync(hbuf.data(), d_totals, (size_t)NT * 4, cudaMemcpyDeviceToHost, stream_));
      CK(cudaStreamSynchronize(stream_));
      CK(cudaGetLastError());
      for (uint32_t k = 0; k < GK_CNT_EXTRA; ++k) htot[k] = hbuf[k];
      for (uint32_t k = 0; k < xh.nbytecols; ++k) btot[k] = hbuf[GK_CNT_EXTRA + NS + k];
    }
    // ---- destination arena (exact sizes) + the pointer tables of the write pass
    Carver ac;
    GkBatch h;
    memset(&h, 0, sizeof h);
    h.n = n;
    h.has_old = 0;
    uint64_t alg = 0;
    auto arr = [&](size_t bytes) {
      alg += bytes;
      return ac.take(bytes);
    };
    const size_t a_flags = arr((size_t)n * 4), a_kind = arr((size_t)n * 4), a_group = arr((size_t)n * 4), a_nsnoff = arr(((size_t)n + 1) * 4),
                 a_nsnb = arr(htot[3]), a_nameoff = arr(((size_t)n + 1) * 4), a_nameb = arr(htot[0]), a_genoff = arr(((size_t)n + 1) * 4),
                 a_genb = arr(htot[1]), a_lbloff = arr(((size_t)n + 1) * 4), a_lblkv = arr((size_t)htot[2] * 8), a_nsrow = arr((size_t)n * 4),
                 a_nsloff = arr(ns.nsl_off.size() * 4), a_nslkv = arr(ns.nsl_kv.size() * 4);
    std::vector<size_t> a_scope(NS, 0);
    for (uint32_t s2 = 1; s2 < NS; ++s2) a_scope[s2] = arr(((size_t)total[xh.scopes[s2].parent] + 1) * 4);
    struct ColOff {
      size_t vt = 0, sid = 0, num = 0, boff = 0, bytes = 0, head = 0;
    };
    std::vector<ColOff> a_col(NC);
    for (uint32_t ci = 0; ci < NC; ++ci) {
      const GkXCol& xc = xh.cols[ci];
      const size_t rows = total[xc.scope];
      if (xc.enc & GK_ENC_VT) a_col[ci].vt = arr(rows);
      if (xc.enc & GK_ENC_SID) a_col[ci].sid = arr(rows * 4);
      if (xc.enc & GK_ENC_NUM) a_col[ci].num = arr(rows * 8);
      if (xc.enc & GK_ENC_HEAD) a_col[ci].head = arr(rows * 32);
      if (xc.enc & GK_ENC_BYTES) {
        a_col[ci].boff = arr((rows + 1) * 4);
        a_col[ci].bytes = arr(btot[xc.bytes_slot]);
      }
    }
    // in-arena tables: GkColumn[], GkScope[], and the pointer arrays of GkIngestOut
    const size_t a_cols = ac.take((size_t)NC * sizeof(GkColumn)), a_scopes = ac.take((size_t)NS * sizeof(GkScope)), a_pscope = ac.take((size_t)NS * 8),
                 a_pvt = ac.take((size_t)NC * 8), a_psid = ac.take((size_t)NC * 8), a_pnum = ac.take((size_t)NC * 8), a_pboff = ac.take((size_t)NC * 8),
                 a_pbytes = ac.take((size_t)NC * 8), a_phead = ac.take((size_t)NC * 8);
    const size_t tables_lo = gk_align(a_cols) == a_cols ? a_cols : a_cols;
    auto* db = new DevBatch();
    std::unique_ptr<DevBatch, std::function<void(DevBatch*)>> guard(db, [this](DevBatch* x) { release(x); });
    db->bytes = gk_align(ac.off);
    db->n = n;
    db->prog_version = c.version;
    dmalloc(&db->arena, db->bytes);
    uint8_t* A = db->arena;
    std::vector<uint8_t> aimg(db->bytes - tables_lo, 0);   // host image of the table tail of the arena
    auto tail = [&](size_t off) { return aimg.data() + (off - tables_lo); };
    GkColumn* hc = reinterpret_cast<GkColumn*>(tail(a_cols));
    GkScope* hs = reinterpret_cast<GkScope*>(tail(a_scopes));
    auto dptr = [&](size_t off) { return reinterpret_cast<uint64_t>(A + off); };
    for (uint32_t ci = 0; ci < NC; ++ci) {
      const GkXCol& xc = xh.cols[ci];
      GkColumn& g = hc[ci];
      g.scope = (int32_t)xc.scope;
      g.enc = xc.enc;
      g.vt = (xc.enc & GK_ENC_VT) ? A + a_col[ci].vt : nullptr;
      g.sid = (xc.enc & GK_ENC_SID) ? reinterpret_cast<const uint32_t*>(A + a_col[ci].sid) : nullptr;
      g.num = (xc.enc & GK_ENC_NUM) ? reinterpret_cast<const int64_t*>(A + a_col[ci].num) : nullptr;
      g.boff = (xc.enc & GK_ENC_BYTES) ? reinterpret_cast<const uint32_t*>(A + a_col[ci].boff) : nullptr;
      g.bytes = (xc.enc & GK_ENC_BYTES) ? A + a_col[ci].bytes : nullptr;
      g.head = (xc.enc & GK_ENC_HEAD) ? reinterpret_cast<const uint32_t*>(A + a_col[ci].head) : nullptr;
      reinterpret_cast<uint64_t*>(tail(a_pvt))[ci] = g.vt ? dptr(a_col[ci].vt) : 0;
      reinterpret_cast<uint64_t*>(tail(a_psid))[ci] = g.sid ? dptr(a_col[ci].sid) : 0;
      reinterpret_cast<uint64_t*>(tail(a_pnum))[ci] = g.num ? dptr(a_col[ci].num) : 0;
      reinterpret_cast<uint64_t*>(tail(a_pboff))[ci] = g.boff ? dptr(a_col[ci].boff) : 0;
      reinterpret_cast<uint64_t*>(tail(a_pbytes))[ci] = g.bytes ? dptr(a_col[ci].bytes) : 0;
      reinterpret_cast<uint64_t*>(tail(a_phead))[ci] = g.head ? dptr(a_col[ci].head) : 0;
    }
    for (uint32_t s2 = 0; s2 < NS; ++s2) {
      hs[s2].parent = s2 ? xh.scopes[s2].parent : 0;
      hs[s2].rows = s2 ? total[s2] : n;
      hs[s2].off = s2 ? reinterpret_cast<const uint32_t*>(A + a_scope[s2]) : nullptr;
      reinterpret_cast<uint64_t*>(tail(a_pscope))[s2] = s2 ? dptr(a_scope[s2]) : 0;
    }
    push_small(A + tables_lo, aimg.data(), aimg.size());
    push_small(A + a_nsloff, ns.nsl_off.data(), ns.nsl_off.size() * 4);
    push_small(A + a_nslkv, ns.nsl_kv.data(), ns.nsl_kv.size() * 4);
    h.flags = reinterpret_cast<const uint32_t*>(A + a_flags);
    h.kind_sid = reinterpret_cast<const uint32_t*>(A + a_kind);
    h.group_sid = reinterpret_cast<const uint32_t*>(A + a_group);
    h.nsn_off = reinterpret_cast<const uint32_t*>(A + a_nsnoff);
    h.nsn_bytes = A + a_nsnb;
    h.name_off = reinterpret_cast<const uint32_t*>(A + a_nameoff);
    h.name_bytes = A + a_nameb;
    h.gen_off = reinterpret_cast<const uint32_t*>(A + a_genoff);
    h.gen_bytes = A + a_genb;
    h.lbl_off = reinterpret_cast<const uint32_t*>(A + a_lbloff);
    h.lbl_kv = reinterpret_cast<const uint32_t*>(A + a_lblkv);
    h.nsrow = reinterpret_cast<const uint32_t*>(A + a_nsrow);
    h.nsl_off = reinterpret_cast<const uint32_t*>(A + a_nsloff);
    h.nsl_kv = reinterpret_cast<const uint32_t*>(A + a_nslkv);
    h.cols = reinterpret_cast<const GkColumn*>(A + a_cols);
    h.scopes = reinterpret_cast<const GkScope*>(A + a_scopes);
    h.ncols = NC;
    h.nscopes = NS;
    db->hdr = h;
    GkIngestOut out;
    memset(&out, 0, sizeof out);
    out.flags = reinterpret_cast<uint32_t*>(A + a_flags);
    out.kind_sid = reinterpret_cast<uint32_t*>(A + a_kind);
    out.group_sid = reinterpret_cast<uint32_t*>(A + a_group);
    out.name_off = reinterpret_cast<uint32_t*>(A + a_nameoff);
    out.name_bytes = A + a_nameb;
    out.gen_off = reinterpret_cast<uint32_t*>(A + a_genoff);
    out.gen_bytes = A + a_genb;
    out.lbl_off = reinterpret_cast<uint32_t*>(A + a_lbloff);
    out.lbl_kv = reinterpret_cast<uint32_t*>(A + a_lblkv);
    out.nsrow = reinterpret_cast<uint32_t*>(A + a_nsrow);
    out.nsn_off = reinterpret_cast<uint32_t*>(A + a_nsnoff);
    out.nsn_bytes = A + a_nsnb;
    out.scope_off = reinterpret_cast<uint32_t* const*>(A + a_pscope);
    out.vt = reinterpret_cast<uint8_t* const*>(A + a_pvt);
    out.sid = reinterpret_cast<uint32_t* const*>(A + a_psid);
    out.num = reinterpret_cast<long long* const*>(A + a_pnum);
    out.boff = reinterpret_cast<uint32_t* const*>(A + a_pboff);
    out.bytes = reinterpret_cast<uint8_t* const*>(A + a_pbytes);
    out.head = reinterpret_cast<uint32_t* const*>(A + a_phead);
    out.row_rec = outc.row_rec;   // (the row handles of the count phase)
    unsigned long long* d_mm = reinterpret_cast<unsigned long long*>(talloc(64));
    {
      const unsigned long long init[2] = {~0ull, 0ull};
      push_small(d_mm, init, 16);
      out.gvk = n ? reinterpret_cast<unsigned long long*>(talloc((size_t)n * 8)) : nullptr;
    }
    if (!n) CK(cudaMemsetAsync(A, 0, db->bytes, stream_));   // (no kernel writes the closing CSR entries of an empty batch)
    // the CSR offsets of the scopes and of the byte columns are the scanned counts of the count phase
    for (uint32_t s2 = 1; s2 < NS; ++s2)
      if (d_cnt[s2]) CK(cudaMemcpyAsync(A + a_scope[s2], d_cnt[s2], ((size_t)total[xh.scopes[s2].parent] + 1) * 4, cudaMemcpyDeviceToDevice, stream_));
    for (uint32_t ci = 0; ci < NC; ++ci) {
      const GkXCol& xc = xh.cols[ci];
      if ((xc.enc & GK_ENC_BYTES) && d_blen[xc.bytes_slot])
        CK(cudaMemcpyAsync(A + a_col[ci].boff, d_blen[xc.bytes_slot], ((size_t)total[xc.scope] + 1) * 4, cudaMemcpyDeviceToDevice, stream_));
    }
    // ---- write passes, repeated while lookups are missing (the host evaluates each distinct argument tuple once)
    CK(cudaEventRecord(e2, stream_));
    uint64_t total_miss = 0;
    double lut_ms = 0;
    std::vector<GkMiss> hmiss;
    for (int round = 0; n && round < 4096; ++round) {
      xp.lut_tab.keys = d_lut_keys_;
      xp.lut_tab.vals = d_lut_vals_;
      xp.lut_tab.mask = lut_.mask;
      xp.lut_vals = d_lutv_;
      CK(cudaMemsetAsync(in.nmiss, 0, 4, stream_));
      if (round == 0 || xh.nbytecols) {   // (the header does not depend on the lookups; a byte column's sid may)
        gk_header_kernel<<<blocks(n), kIngestThreads, 0, stream_>>>(xp, in, out);
        ++launches_;
        for (uint32_t ci = 0; ci < NC; ++ci) {
          const GkXCol& xc = xh.cols[ci];
          if (!(xc.enc & GK_ENC_BYTES) || !total[xc.scope]) continue;
          gk_bcol_write_kernel<<<blocks(total[xc.scope]), kIngestThreads, 0, stream_>>>(xp, in, out, ci, total[xc.scope]);
          ++launches_;
        }
      }
      for (uint32_t s2 = 0; s2 < NS; ++s2) {
        const uint32_t rows = total[s2];
        if (!rows || !xh.scopes[s2].ncols) continue;
        gk_cols_kernel<<<(uint32_t)(((uint64_t)rows * clanes + kIngestThreads - 1) / kIngestThreads), kIngestThreads, 0, stream_>>>(xp, in, out, s2, rows, clanes);
        ++launches_;
      }
      uint32_t nm = 0;
      CK(cudaMemcpyAsync(&nm, in.nmiss, 4, cudaMemcpyDeviceToHost, stream_));
      CK(cudaStreamSynchronize(stream_));
      CK(cudaGetLastError());
      if (nm == 0) break;
      const auto L0 = std::chrono::steady_clock::now();
      const uint32_t m = std::min(nm, miss_cap);
      hmiss.resize(m);
      CK(cudaMemcpy(hmiss.data(), in.misses, (size_t)m * sizeof(GkMiss), cudaMemcpyDeviceToHost));
      std::vector<GkLutVal> vals;
      rq.lut_fill(hmiss.data(), m, vals);
      // append the results, point the claimed slots at them
      const uint32_t base = (uint32_t)lutv_.size();
      lutv_.insert(lutv_.end(), vals.begin(), vals.end());
      if (lutv_.size() > d_lutv_cap_) {
        GkLutVal* nv = nullptr;
        const size_t ncap = lutv_.size() * 2 + 1024;
        CK(cudaMalloc(&nv, ncap * sizeof(GkLutVal)));
        if (base) CK(cudaMemcpy(nv, d_lutv_, (size_t)base * sizeof(GkLutVal), cudaMemcpyDeviceToDevice));
        if (d_lutv_) cudaFree(d_lutv_);
        d_lutv_ = nv;
        d_lutv_cap_ = ncap;
      }
      CK(cudaMemcpy(d_lutv_ + base, vals.data(), (size_t)m * sizeof(GkLutVal), cudaMemcpyHostToDevice));
      std::vector<uint32_t> pairs((size_t)m * 2);
      for (uint32_t j = 0; j < m; ++j) {
        pairs[2 * j] = hmiss[j].slot;
        pairs[2 * j + 1] = base + j;
        lut_filled_.emplace_back(hmiss[j].key, base + j);
      }
      CK(cudaMemcpy(d_fill, pairs.data(), pairs.size() * 4, cudaMemcpyHostToDevice));
      gk_fill_kernel<<<(m + 255) / 256, 256, 0, stream_>>>(d_lut_vals_, d_fill, m);
      ++launches_;
      total_miss += m;
      if ((lut_filled_.size() + (nm - m)) * 2 > (size_t)lut_.mask) {   // grow: filled entries re-inserted, pending claims dropped
        CK(cudaStreamSynchronize(stream_));
        uint32_t cap = (lut_.mask + 1) * 4;
        while ((size_t)cap < lut_filled_.size() * 4) cap <<= 1;
        lut_.init(cap);
        for (auto& kv : lut_filled_) lut_.put(kv.first, kv.second);
        upload_htab(lut_, &d_lut_keys_, &d_lut_vals_, &d_lut_cap_);
      }
      lut_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - L0).count();
      if (round == 4095) throw BackendError{"device ingest: lookups did not converge"};
    }
    CK(cudaEventRecord(e3, stream_));
    // ---- status of every object, tiling of the evaluation kernel
    if (status) {
      status->assign(n, 0);
      if (n) CK(cudaMemcpyAsync(status->data(), in.status, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_));
    }
    uint32_t tile = kTile;
    if (const char* ft = getenv("GK_FORCE_TILE")) tile = std::min<uint32_t>(kTile, std::max(32, atoi(ft)) / 32 * 32);
    unsigned long long mm[2] = {~0ull, 0ull};
    if (n) {
      gk_gvk_minmax_kernel<<<std::min<uint32_t>(1024u, (n + 255u) / 256u), 256, 0, stream_>>>(out.gvk, n, d_mm);
      ++launches_;
      CK(cudaMemcpyAsync(mm, d_mm, 16, cudaMemcpyDeviceToHost, stream_));
    }
    std::vector<uint32_t> cap(NS, 0);
    uint32_t ntiles = 0;
    for (;; tile = std::max(32u, tile / 2 / 32 * 32)) {   // (halved until the tile's bit columns fit the CTA's shared memory)
      ntiles = (n + tile - 1) / tile;
      if (db->d_tile_lo) dfree(db->d_tile_lo), db->d_tile_lo = nullptr;
      dmalloc(&db->d_tile_lo, ((size_t)(ntiles + 1) * NS) * 4 + 64);
      CK(cudaMemsetAsync(d_cap, 0, (size_t)NS * 4, stream_));
      gk_tiles_kernel<<<(ntiles + 1 + 127) / 128, 128, 0, stream_>>>(xp, out, n, NS, tile, ntiles, db->d_tile_lo, d_cap);
      ++launches_;
      CK(cudaMemcpyAsync(cap.data(), d_cap, (size_t)NS * 4, cudaMemcpyDeviceToHost, stream_));
      CK(cudaStreamSynchronize(stream_));
      CK(cudaGetLastError());
      if (tile <= 32 || tile_fits(c, cap, tile)) break;
    }
    db->tile = tile;
    db->ntiles = ntiles;
    db->gvk_uniform = mm[1] == 0 || mm[0] == mm[1];
    finish_batch(db, c, cap);
    if (st) {
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      st->h2d_ms = ms;   // copy + tokenise, overlapped chunk by chunk
      st->tape_ms = ms;
      cudaEventElapsedTime(&ms, e1, e3);
      st->extract_ms = ms - lut_ms;
      st->lut_ms = lut_ms;
      st->h2d_bytes = B + ((size_t)n + 1) * 8 + tab_bytes + aimg.size();
      st->lut_misses = total_miss;
      st->alg_bytes = alg;
      st->launches = launches_;
      st->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count();
    }
    if (getenv("GK_TRACE_INGEST")) {
      float a = 0, b = 0, d = 0;
      cudaEventElapsedTime(&a, e0, e1);
      cudaEventElapsedTime(&b, e1, e2);
      cudaEventElapsedTime(&d, e2, e3);
      fprintf(stderr, "[ingest] n=%u blob %.1f MB: copy+tokenise %.2f ms, count+scan(+alloc) %.2f ms, write(+lookups: %llu missed, %.1f ms host) %.2f ms, total wall %.2f ms, arena %.1f MB\n",
              n, B / 1e6, a, b, (unsigned long long)total_miss, lut_ms, d,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count(), db->bytes / 1e6);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    cudaEventDestroy(e3);
    guard.release();
    return db;
  }
  void* fork_batch(void* b, const Compiled& other) override {
    auto* db = static_cast<DevBatch*>(b);
    CK(cudaSetDevice(device_));
    auto* f = new DevBatch();
    f->arena = db->arena;
    f->bytes = db->bytes;
    f->hdr = db->hdr;
    f->n = db->n;
    f->d_tile_lo = db->d_tile_lo;
    f->ntiles = db->ntiles;
    f->tile = db->tile;
    f->gvk_uniform = db->gvk_uniform;
    f->fork = true;
    f->prog_version = other.version;
    std::unique_ptr<DevBatch, std::function<void(DevBatch*)>> guard(f, [this](DevBatch* x) { release(x); });
    finish_batch(f, other, db->cap);
    guard.release();
    return f;
  }

  void identity(void* b, BatchIdentity& out) override {
    auto* db = static_cast<DevBatch*>(b);
    std::lock_guard<std::mutex> l(mu_);
    CK(cudaSetDevice(device_));
    const uint32_t n = db->n;
    out.uniform_gvk = db->gvk_uniform;
    out.flags.resize(n);
    out.ns_off.resize((size_t)n + 1);
    out.name_off.resize((size_t)n + 1);
    if (!n) {
      out.ns_off[0] = out.name_off[0] = 0;
      return;
    }
    CK(cudaMemcpyAsync(out.flags.data(), db->hdr.flags, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(out.ns_off.data(), db->hdr.nsn_off, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(out.name_off.data(), db->hdr.name_off, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    out.ns_bytes.resize(out.ns_off[n]);
    out.name_bytes.resize(out.name_off[n]);
    if (!out.ns_bytes.empty()) CK(cudaMemcpyAsync(out.ns_bytes.data(), db->hdr.nsn_bytes, out.ns_bytes.size(), cudaMemcpyDeviceToHost, stream_));
    if (!out.name_bytes.empty()) CK(cudaMemcpyAsync(out.name_bytes.data(), db->hdr.name_bytes, out.name_bytes.size(), cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
  }

  // ---- front buffers: a page's JSON, offsets and tape.  Two of them, so that page k+1 streams in (copy stream + front stream)
  // while page k is extracted and evaluated on the main stream.
  struct Front {
    Scratch buf;
    GkIngestIn in{};            // blob / ooff / tape / ntape / status pointers into buf, n
    const uint8_t* host_blob = nullptr;
    size_t n = 0;
    bool pending = false;       // prefetched, not yet consumed
    cudaEvent_t done = nullptr; // copy + tokenise finished
  };
  Front& start_front(const uint8_t* blob, const unsigned long long* ooff, size_t nn) {   // (front_mu_ held)
    Front& f = fronts_[next_front_];
    next_front_ ^= 1;
    const uint32_t n = (uint32_t)nn;
    const unsigned long long B = ooff[n];
    Carver sc;
    const size_t o_blob = sc.take((size_t)B + 16), o_ooff = sc.take(((size_t)n + 1) * 8), o_tape = sc.take(((size_t)(B / 2) + 4ull * n + 64) * 8),
                 o_ntape = sc.take((size_t)n * 4), o_status = sc.take((size_t)n * 4);
    if (!f.done) CK(cudaEventCreateWithFlags(&f.done, cudaEventDisableTiming));
    // the buffer may still be read by the kernels of the page that used it last: they ran on the main stream
    CK(cudaStreamSynchronize(stream_));
    uint8_t* d = f.buf.need(gk_align(sc.off));
    memset(&f.in, 0, sizeof f.in);
    f.in.blob = d + o_blob;
    f.in.ooff = reinterpret_cast<const unsigned long long*>(d + o_ooff);
    f.in.tape = reinterpret_cast<unsigned long long*>(d + o_tape);
    f.in.ntape = reinterpret_cast<uint32_t*>(d + o_ntape);
    f.in.status = reinterpret_cast<uint32_t*>(d + o_status);
    f.in.n = n;
    f.host_blob = blob;
    f.n = nn;
    // H2D of the raw JSON in chunks on the copy stream; the tokeniser of a chunk starts as soon as its bytes have landed
    CK(cudaMemcpyAsync(d + o_ooff, ooff, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, copy_stream_));
    // Chunk size: a tokeniser launch of a 32 MB chunk (50k Pods) fills 17 % of the warp slots and takes ~0.6 ms whatever its size
    // (one thread walks one object); bigger chunks mean fewer, fuller launches at the price of a later start.  GK_INGEST_CHUNK_MB
    // overrides (8 MB chunks measured 85 ms per page instead of 19).
    static const size_t kChunk = []() {
      size_t mb = 32;
      if (const char* ev = getenv("GK_INGEST_CHUNK_MB")) mb = (size_t)std::max(1, atoi(ev));
      return mb << 20;
    }();
    uint32_t first = 0;
    while (first < n) {
      uint32_t last = first;
      const unsigned long long lo = ooff[first];
      while (last < n && ooff[last + 1] - lo <= kChunk) ++last;
      if (last == first) ++last;   // one object larger than a chunk
      const unsigned long long hi = ooff[last];
      CK(cudaMemcpyAsync(d + o_blob + lo, blob + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, copy_stream_));
      cudaEvent_t ev = chunk_event(first);
      CK(cudaEventRecord(ev, copy_stream_));
      CK(cudaStreamWaitEvent(front_stream_, ev, 0));
      const uint32_t cnt = last - first;
      gk_tape_kernel<<<(cnt + kIngestThreads - 1) / kIngestThreads, kIngestThreads, 0, front_stream_>>>(f.in, first, cnt);
      ++launches_;
      first = last;
    }
    if (n == 0) CK(cudaStreamWaitEvent(front_stream_, chunk_event_after_copy(), 0));
    CK(cudaEventRecord(f.done, front_stream_));
    f.pending = true;
    return f;
  }
  cudaEvent_t chunk_event_after_copy() {
    cudaEvent_t ev = chunk_event(0);
    CK(cudaEventRecord(ev, copy_stream_));
    return ev;
  }
  void prefetch(const uint8_t* blob, const unsigned long long* ooff, size_t n) override {
    std::lock_guard<std::mutex> fl(front_mu_);
    CK(cudaSetDevice(device_));
    start_front(blob, ooff, n);
  }

  void pin_host(const void* p, size_t bytes, bool pin) override {
    CK(cudaSetDevice(device_));
    if (pin) CK(cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault));
    else CK(cudaHostUnregister(const_cast<void*>(p)));
  }
  cudaEvent_t chunk_event(uint32_t k) {
    (void)k;
    if (chunk_ev_next_ >= chunk_evs_.size()) {
      cudaEvent_t e;
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      chunk_evs_.push_back(e);
    }
    cudaEvent_t e = chunk_evs_[chunk_ev_next_];
    chunk_ev_next_ = (chunk_ev_next_ + 1) % 64 == 0 ? 0 : chunk_ev_next_ + 1;
    return e;
  }

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
  std::mutex ingest_mu_;
  Scratch scratch_, tabs_, rows_;
  SidTable sid_;
  unsigned long long* d_sid_keys_ = nullptr;
  uint32_t* d_sid_vals_ = nullptr;
  size_t d_sid_cap_ = 0;
  HashTabHost lut_;                                               // geometry of the device table (the device copy is authoritative)
  std::vector<std::pair<unsigned long long, uint32_t>> lut_filled_;   // every answered lookup (key, result index): re-inserted on growth
  unsigned long long* d_lut_keys_ = nullptr;
  uint32_t* d_lut_vals_ = nullptr;
  size_t d_lut_cap_ = 0;
  std::vector<GkLutVal> lutv_;
  GkLutVal* d_lutv_ = nullptr;
  size_t d_lutv_cap_ = 0;
  cudaStream_t copy_stream_ = nullptr, front_stream_ = nullptr;
  std::mutex front_mu_;
  Front fronts_[2];
  int next_front_ = 0;
  std::vector<cudaEvent_t> chunk_evs_;
  size_t chunk_ev_next_ = 0;
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
  std::list<SpecEntry> spec_cache_;
  SpecEntry* spec_cur_ = nullptr;
  bool spec_on_ = !(getenv("GK_SPEC") && atoi(getenv("GK_SPEC")) == 0);
  uint32_t spec_min_objects_ = getenv("GK_SPEC_MIN_OBJECTS") ? (uint32_t)atoll(getenv("GK_SPEC_MIN_OBJECTS")) : 8192u;   // below: not worth a ~2 s NVRTC run
  // one thread per object of a 512-object tile, 128 registers: the fastest of the shapes measured (profiles/experiments/README.md)
  int spec_threads_ = getenv("GK_SPEC_THREADS") ? std::min(1024, std::max(32, atoi(getenv("GK_SPEC_THREADS")) / 32 * 32)) : 512;
  int spec_minb_ = getenv("GK_SPEC_MINB") ? std::max(1, atoi(getenv("GK_SPEC_MINB"))) : 1;
  const char* last_kernel_ = "gk_eval_kernel";
  bool spec_next_ = false;
  void* pinned_ = nullptr;
  size_t pinned_bytes_ = 0;
};

Backend* make_backend(int device) { return new CudaBackend(device); }

}  // namespace gk
