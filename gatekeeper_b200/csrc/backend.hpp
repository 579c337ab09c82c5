// Execution backend interface.  The product library links exactly one implementation: the CUDA backend
// (kernels.cu).  tests/_hostemu links the CPU emulation of the same per-object core (vm_core.h) so the
// lowering/flattening logic can be unit-tested in a container without a GPU; it is never part of the
// product library and the product has no CPU fallback.
#pragma once
#include <atomic>
#include <thread>
#include <cstdint>
#include <cstring>
#include <new>
#include <utility>
#include <string>
#include <vector>

#include "engine.hpp"
#include "program.h"

namespace gk {

struct BackendError {
  std::string msg;
};

// Result bitmaps land in page-locked host memory when the backend has any (the CUDA backend installs the hooks): a device->host
// copy into pageable memory runs at a fraction of the link rate.  Blocks are recycled through a small size-bucketed pool
// (page-locking is expensive: ~1 ms per 8 MB).
struct HostBlockHooks {
  void* (*alloc)(size_t) = nullptr;   // null: plain malloc / free
  void (*release)(void*) = nullptr;
};
HostBlockHooks& host_block_hooks();
void* host_block_take(size_t bytes);
void host_block_give(void* p, size_t bytes);
template <class T>
struct HostBlockAlloc {
  using value_type = T;
  HostBlockAlloc() = default;
  template <class U>
  HostBlockAlloc(const HostBlockAlloc<U>&) {}
  T* allocate(size_t n) { return static_cast<T*>(host_block_take(n * sizeof(T))); }
  void deallocate(T* p, size_t n) { host_block_give(p, n * sizeof(T)); }
  template <class U>
  void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }   // resize() leaves the words unset: the copy-back overwrites all of them
  template <class U, class... A>
  void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
  template <class U>
  bool operator==(const HostBlockAlloc<U>&) const { return true; }
  template <class U>
  bool operator!=(const HostBlockAlloc<U>&) const { return false; }
};
using BitPlane = std::vector<uint32_t, HostBlockAlloc<uint32_t>>;

struct EvalOut {
  uint32_t n = 0, nconstraints = 0, words = 0;
  BitPlane viol, err;                       // [n * words]  (filled when copy_back)
  std::vector<uint64_t> totals, err_totals; // [nconstraints]
  std::vector<uint32_t> errlist;            // triples (object, constraint, code)
  float kernel_ms = 0.f;
  uint64_t launches = 0;
};

struct DevOutPtrs {          // caller-owned device buffers (e.g. torch tensors) -- used by the multi-GPU path
  void* viol = nullptr;      // u32 [n * words]
  void* err = nullptr;       // u32 [n * words]
  void* totals = nullptr;    // u64 [nconstraints]
  void* err_totals = nullptr;
  void* stream = nullptr;    // cudaStream_t
  // fused exchange: device addresses, on every peer GPU, of the place THIS rank's bitmap shard / totals go
  uint64_t peer_viol[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t peer_tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t npeers = 0, tot_stride = 0;
};

// What orders audit results besides the message: group / version / kind, namespace, name (pkg/audit/manager.go:161-202).  Read
// back from the batch's header arrays; `uniform_gvk`: every object that was not skipped has the same apiVersion and kind.
struct BatchIdentity {
  bool uniform_gvk = false;
  std::vector<uint32_t> flags, ns_off, name_off;
  std::vector<uint8_t> ns_bytes, name_bytes;
};

class Backend {
 public:
  virtual ~Backend() {}
  virtual const char* name() const = 0;
  // which kernel decided the last evaluation: "gk_spec_kernel" (generated for the constraint set, spec_codegen.hpp) or "gk_eval_kernel"
  virtual const char* last_kernel() const { return "gk_eval_kernel"; }
  virtual void set_program(const Compiled& c) = 0;                       // upload tables when the version changed
  virtual void sync_strings(const StringTable& st) = 0;                  // (re)upload the dictionary if it grew
  virtual void* upload(const HostBatch& hb, const Compiled& c, double* h2d_ms, uint64_t* h2d_bytes) = 0;
  virtual void release(void* batch) = 0;
  virtual void eval(void* batch, const std::vector<uint32_t>& active, EvalOut& out, bool copy_back) = 0;
  virtual void eval_into(void* batch, const std::vector<uint32_t>& active, const DevOutPtrs& dst) = 0;
  // Device ingest (ingest_core.h): raw JSON blob -> resident columnar batch without a host parse.  `status` receives one
  // GK_ING_* code per object (the caller renders the error text of the rare non-OK ones with the host parser).
  virtual void* ingest(const IngestReq& rq, IngestStats* st, std::vector<uint32_t>* status) = 0;
  virtual void identity(void* batch, BatchIdentity& out) = 0;
  // A second view of a resident batch for ANOTHER netlist over the same schema (Compiled::amb): shares the columns, owns its
  // own slot-resolved netlist and output planes.  Evaluate it with eval() while the backend holds `other`; release() it before
  // the batch it was forked from.
  virtual void* fork_batch(void* batch, const Compiled& other) = 0;
  // Start moving a blob to the device ahead of its ingest() (copy + tokenise on the copy / front streams, into the idle one of
  // two front buffers) and return at once: the next page of an audit sweep streams in while the current one is being extracted
  // and evaluated.  ingest() of the same (blob, n) picks the prefetched copy up; anything else is ingested from scratch.
  virtual void prefetch(const uint8_t* blob, const unsigned long long* ooff, size_t n) { (void)blob, (void)ooff, (void)n; }
  // page-lock (or release) a caller-owned host buffer so that the blob copy is a direct DMA at link speed
  virtual void pin_host(const void* p, size_t bytes, bool pin) { (void)p, (void)bytes, (void)pin; }
};

Backend* make_backend(int device);   // defined by whichever backend the library links

// ---- arena packing shared by both backends: every array of a HostBatch laid out contiguously
struct PackedBatch {
  std::vector<uint8_t> arena;          // host image
  size_t cols_off = 0, scopes_off = 0; // where the GkColumn[] / GkScope[] tables live in the arena
  GkBatch hdr{};                       // pointers are OFFSETS into the arena until rebased
};

inline size_t gk_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Packing is two steps so that the bytes are copied exactly once, by all host threads, straight into the buffer the
// device copy reads from (pinned memory for the CUDA backend):
//   pack_layout : assigns every array its 256-byte aligned offset in the arena image and records one copy job per array
//   pack_copy   : runs the copy jobs (memcpy, parallel over arrays and over slices of big arrays) into `dst`
struct PackJob {
  const void* src;
  size_t bytes, off;
};
struct PackPlan {
  std::vector<PackJob> jobs;
  std::vector<GkColumn> cols;      // tables that live in the arena themselves (offsets, rebased later)
  std::vector<GkScope> scopes;
  size_t total = 0;
};

template <class T>
inline size_t plan_put(PackPlan& pl, const std::vector<T>& v) {
  size_t off = gk_align(pl.total);
  pl.total = off + std::max<size_t>(v.size() * sizeof(T), 16);
  if (!v.empty()) pl.jobs.push_back({v.data(), v.size() * sizeof(T), off});
  return off;
}

inline void pack_layout(const HostBatch& hb, const Compiled& c, PackedBatch& pb, PackPlan& pl) {
  GkBatch& h = pb.hdr;
  memset(&h, 0, sizeof h);
  h.n = hb.n;
  h.has_old = hb.has_old ? 1 : 0;
#define PUT(field, vec) h.field = reinterpret_cast<decltype(h.field)>(plan_put(pl, vec))
  PUT(flags, hb.flags);
  PUT(kind_sid, hb.kind_sid);
  PUT(group_sid, hb.group_sid);
  PUT(nsn_off, hb.nsn_off);
  PUT(nsn_bytes, hb.nsn_bytes);
  PUT(name_off, hb.name_off);
  PUT(name_bytes, hb.name_bytes);
  PUT(gen_off, hb.gen_off);
  PUT(gen_bytes, hb.gen_bytes);
  PUT(lbl_off, hb.lbl_off);
  PUT(lbl_kv, hb.lbl_kv);
  PUT(nsrow, hb.nsrow);
  PUT(nsl_off, hb.nsl_off);
  PUT(nsl_kv, hb.nsl_kv);
#undef PUT
  pl.cols.resize(hb.cols.size());
  for (size_t i = 0; i < hb.cols.size(); ++i) {
    GkColumn& g = pl.cols[i];
    memset(&g, 0, sizeof g);
    g.scope = c.schema.cols[i].scope;
    g.enc = c.schema.cols[i].enc;
    g.vt = reinterpret_cast<const uint8_t*>(plan_put(pl, hb.cols[i].vt));
    g.sid = reinterpret_cast<const uint32_t*>(plan_put(pl, hb.cols[i].sid));
    g.num = reinterpret_cast<const int64_t*>(plan_put(pl, hb.cols[i].num));
    g.boff = reinterpret_cast<const uint32_t*>(plan_put(pl, hb.cols[i].boff));
    g.bytes = reinterpret_cast<const uint8_t*>(plan_put(pl, hb.cols[i].bytes));
    g.head = reinterpret_cast<const uint32_t*>(plan_put(pl, hb.cols[i].head));
  }
  pl.scopes.resize(hb.scope_off.size());
  for (size_t s = 0; s < pl.scopes.size(); ++s) {
    memset(&pl.scopes[s], 0, sizeof(GkScope));
    pl.scopes[s].parent = c.schema.scopes[s].parent;
    pl.scopes[s].rows = hb.scope_rows[s];
    pl.scopes[s].off = reinterpret_cast<const uint32_t*>(plan_put(pl, hb.scope_off[s]));
  }
  pb.cols_off = plan_put(pl, pl.cols);
  pb.scopes_off = plan_put(pl, pl.scopes);
  h.cols = reinterpret_cast<const GkColumn*>(pb.cols_off);
  h.scopes = reinterpret_cast<const GkScope*>(pb.scopes_off);
  h.ncols = (uint32_t)pl.cols.size();
  h.nscopes = (uint32_t)pl.scopes.size();
  pl.total = gk_align(pl.total);
}

inline void pack_copy(const PackPlan& pl, uint8_t* dst, int threads) {
  // slice big arrays so that the work spreads evenly
  struct Piece {
    const uint8_t* src;
    uint8_t* dst;
    size_t bytes;
  };
  std::vector<Piece> pieces;
  const size_t kSlice = 4u << 20;
  for (auto& j : pl.jobs)
    for (size_t o = 0; o < j.bytes; o += kSlice)
      pieces.push_back({static_cast<const uint8_t*>(j.src) + o, dst + j.off + o, std::min(kSlice, j.bytes - o)});
  const size_t T = std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, threads), pieces.size() / 4 + 1));
  if (T == 1) {
    for (auto& p : pieces) memcpy(p.dst, p.src, p.bytes);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> th;
  for (size_t t = 0; t < T; ++t)
    th.emplace_back([&]() {
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= pieces.size()) break;
        memcpy(pieces[i].dst, pieces[i].src, pieces[i].bytes);
      }
    });
  for (auto& x : th) x.join();
}

// one-shot form (host image in pb.arena): used by the test backend
inline void pack_batch(const HostBatch& hb, const Compiled& c, PackedBatch& pb) {
  PackPlan pl;
  pack_layout(hb, c, pb, pl);
  pb.arena.assign(pl.total, 0);
  pack_copy(pl, pb.arena.data(), 1);
}

// turn the arena-relative offsets into pointers valid at `base` (device or host); patches the in-arena tables
inline GkBatch rebase_batch(PackedBatch& pb, uint8_t* host_image, const uint8_t* base) {
  GkBatch h = pb.hdr;
#define RB(f) h.f = reinterpret_cast<decltype(h.f)>(base + reinterpret_cast<size_t>(h.f))
  RB(flags); RB(kind_sid); RB(group_sid); RB(nsn_off); RB(nsn_bytes); RB(name_off); RB(name_bytes); RB(gen_off); RB(gen_bytes);
  RB(lbl_off); RB(lbl_kv); RB(nsrow); RB(nsl_off); RB(nsl_kv); RB(cols); RB(scopes);
#undef RB
  GkColumn* cols = reinterpret_cast<GkColumn*>(host_image + pb.cols_off);
  for (uint32_t i = 0; i < h.ncols; ++i) {
    GkColumn& g = cols[i];
    g.vt = base + reinterpret_cast<size_t>(g.vt);
    g.sid = reinterpret_cast<const uint32_t*>(base + reinterpret_cast<size_t>(g.sid));
    g.num = reinterpret_cast<const int64_t*>(base + reinterpret_cast<size_t>(g.num));
    g.boff = reinterpret_cast<const uint32_t*>(base + reinterpret_cast<size_t>(g.boff));
    g.bytes = base + reinterpret_cast<size_t>(g.bytes);
    g.head = reinterpret_cast<const uint32_t*>(base + reinterpret_cast<size_t>(g.head));
  }
  GkScope* scopes = reinterpret_cast<GkScope*>(host_image + pb.scopes_off);
  for (uint32_t s = 0; s < h.nscopes; ++s) scopes[s].off = reinterpret_cast<const uint32_t*>(base + reinterpret_cast<size_t>(scopes[s].off));
  return h;
}

}  // namespace gk
