// Execution backend interface.  The product library links exactly one implementation: the CUDA backend
// (kernels.cu).  tests/_hostemu links the CPU emulation of the same per-object core (vm_core.h) so the
// lowering/flattening logic can be unit-tested in a container without a GPU; it is never part of the
// product library and the product has no CPU fallback.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "engine.hpp"
#include "program.h"

namespace gk {

struct BackendError {
  std::string msg;
};

struct EvalOut {
  uint32_t n = 0, nconstraints = 0, words = 0;
  std::vector<uint32_t> viol, err;          // [n * words]  (filled when copy_back)
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
};

class Backend {
 public:
  virtual ~Backend() {}
  virtual const char* name() const = 0;
  virtual void set_program(const Compiled& c) = 0;                       // upload tables when the version changed
  virtual void sync_strings(const StringTable& st) = 0;                  // (re)upload the dictionary if it grew
  virtual void* upload(const HostBatch& hb, const Compiled& c, double* h2d_ms, uint64_t* h2d_bytes) = 0;
  virtual void release(void* batch) = 0;
  virtual void eval(void* batch, const std::vector<uint32_t>& active, EvalOut& out, bool copy_back) = 0;
  virtual void eval_into(void* batch, const std::vector<uint32_t>& active, const DevOutPtrs& dst) = 0;
};

Backend* make_backend(int device);   // defined by whichever backend the library links

// ---- arena packing shared by both backends: every array of a HostBatch laid out contiguously
struct PackedBatch {
  std::vector<uint8_t> arena;          // host image
  size_t cols_off = 0, scopes_off = 0; // where the GkColumn[] / GkScope[] tables live in the arena
  GkBatch hdr{};                       // pointers are OFFSETS into the arena until rebased
};

inline size_t gk_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

template <class T>
inline size_t pack_put(std::vector<uint8_t>& arena, const std::vector<T>& v) {
  size_t off = gk_align(arena.size());
  arena.resize(off + std::max<size_t>(v.size() * sizeof(T), 16));
  if (!v.empty()) memcpy(arena.data() + off, v.data(), v.size() * sizeof(T));
  return off;
}

inline void pack_batch(const HostBatch& hb, const Compiled& c, PackedBatch& pb) {
  auto& a = pb.arena;
  a.clear();
  size_t total = 4096;
  total += (hb.flags.size() * 4 + 256) * 4 + hb.name_off.size() * 4 + hb.gen_off.size() * 4 + hb.lbl_off.size() * 4 + hb.lbl_kv.size() * 4 +
           hb.name_bytes.size() + hb.gen_bytes.size() + hb.nsrow.size() * 4 + hb.nsl_off.size() * 4 + hb.nsl_kv.size() * 4 + 4096;
  for (auto& s : hb.scope_off) total += s.size() * 4 + 512;
  for (auto& col : hb.cols) total += col.vt.size() + col.sid.size() * 4 + col.num.size() * 8 + col.boff.size() * 4 + col.bytes.size() + col.head.size() * 4 + 2560;
  total += hb.cols.size() * sizeof(GkColumn) + hb.scope_off.size() * sizeof(GkScope) + 1024;
  a.reserve(total);
  GkBatch& h = pb.hdr;
  memset(&h, 0, sizeof h);
  h.n = hb.n;
  h.has_old = hb.has_old ? 1 : 0;
#define PUT(field, vec) h.field = reinterpret_cast<decltype(h.field)>(pack_put(a, vec))
  PUT(flags, hb.flags);
  PUT(kind_sid, hb.kind_sid);
  PUT(group_sid, hb.group_sid);
  PUT(nsname_sid, hb.nsname_sid);
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
  std::vector<GkColumn> cols(hb.cols.size());
  for (size_t i = 0; i < hb.cols.size(); ++i) {
    GkColumn& g = cols[i];
    memset(&g, 0, sizeof g);
    g.scope = c.schema.cols[i].scope;
    g.enc = c.schema.cols[i].enc;
    g.vt = reinterpret_cast<const uint8_t*>(pack_put(a, hb.cols[i].vt));
    g.sid = reinterpret_cast<const uint32_t*>(pack_put(a, hb.cols[i].sid));
    g.num = reinterpret_cast<const int64_t*>(pack_put(a, hb.cols[i].num));
    g.boff = reinterpret_cast<const uint32_t*>(pack_put(a, hb.cols[i].boff));
    g.bytes = reinterpret_cast<const uint8_t*>(pack_put(a, hb.cols[i].bytes));
    g.head = reinterpret_cast<const uint32_t*>(pack_put(a, hb.cols[i].head));
  }
  std::vector<GkScope> scopes(hb.scope_off.size());
  for (size_t s = 0; s < scopes.size(); ++s) {
    memset(&scopes[s], 0, sizeof(GkScope));
    scopes[s].parent = c.schema.scopes[s].parent;
    scopes[s].rows = hb.scope_rows[s];
    scopes[s].off = reinterpret_cast<const uint32_t*>(pack_put(a, hb.scope_off[s]));
  }
  pb.cols_off = pack_put(a, cols);
  pb.scopes_off = pack_put(a, scopes);
  h.cols = reinterpret_cast<const GkColumn*>(pb.cols_off);
  h.scopes = reinterpret_cast<const GkScope*>(pb.scopes_off);
  h.ncols = (uint32_t)cols.size();
  h.nscopes = (uint32_t)scopes.size();
}

// turn the arena-relative offsets into pointers valid at `base` (device or host); patches the in-arena tables
inline GkBatch rebase_batch(PackedBatch& pb, uint8_t* host_image, const uint8_t* base) {
  GkBatch h = pb.hdr;
#define RB(f) h.f = reinterpret_cast<decltype(h.f)>(base + reinterpret_cast<size_t>(h.f))
  RB(flags); RB(kind_sid); RB(group_sid); RB(nsname_sid); RB(name_off); RB(name_bytes); RB(gen_off); RB(gen_bytes);
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
