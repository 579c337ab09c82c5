// TEST-ONLY backend: executes the constraint netlist (gatekeeper_b200/csrc/program.h GkOp) on the CPU with the
// whole batch as one tile, using the same per-row core (csrc/vm_core.h: gk_atom, gk_match) as the CUDA tile
// executor, so the lowering + flattening logic can be unit-tested in the authoring container (no GPU).
// It is linked ONLY into tests/_hostemu/libgk_hostemu.so.  The product library
// (gatekeeper_b200/libgk_engine.so) links kernels.cu instead and has no CPU path at all.
#include <algorithm>
#include <chrono>
#include <memory>
#include <mutex>

#include "../../gatekeeper_b200/csrc/backend.hpp"
#include "../../gatekeeper_b200/csrc/vm_core.h"

namespace gk {

struct EmuBatch {
  PackedBatch pb;
  GkBatch hdr{};
  uint32_t n = 0, words = 1;
  std::vector<uint32_t> scope_rows;
  std::vector<std::vector<uint32_t>> scope_off;   // host copies for range lookups
};

class HostEmuBackend : public Backend {
 public:
  const char* name() const override { return "hostemu-TEST-ONLY"; }
  void set_program(const Compiled& c) override {
    std::lock_guard<std::mutex> l(mu_);
    prog_ = &c;
  }
  // the dictionary is replaced, never edited in place: an evaluation keeps the copy it started with
  void sync_strings(const StringTable& st) override {
    auto d = std::make_shared<Dict>();
    st.snapshot(d->off, d->bytes);
    std::lock_guard<std::mutex> l(mu_);
    dict_ = std::move(d);
  }
  void* upload(const HostBatch& hb, const Compiled& c, double* ms, uint64_t* bytes) override {
    auto* b = new EmuBatch();
    pack_batch(hb, c, b->pb);
    b->hdr = rebase_batch(b->pb, b->pb.arena.data(), b->pb.arena.data());
    b->n = hb.n;
    b->words = std::max<uint32_t>(1, (uint32_t)((c.cons_match.size() + 31) / 32));
    b->scope_rows = hb.scope_rows;
    b->scope_off = hb.scope_off;
    if (ms) *ms = 0;
    if (bytes) *bytes = b->pb.arena.size();
    return b;
  }
  void release(void* b) override { delete static_cast<EmuBatch*>(b); }

  void eval(void* bb, const std::vector<uint32_t>& active, EvalOut& out, bool) override {
    auto* b = static_cast<EmuBatch*>(bb);
    std::shared_ptr<const Dict> dict;
    const Compiled* progp;
    {
      std::lock_guard<std::mutex> l(mu_);
      dict = dict_;
      progp = prog_;
    }
    const Compiled& c = *progp;
    const uint32_t C = (uint32_t)c.cons_match.size(), W = b->words, n = b->n;
    GkBatch h = b->hdr;
    h.dict_off = dict->off.data();
    h.dict_bytes = dict->bytes.data();
    h.dict_n = (uint32_t)dict->off.size() - 1;
    out.n = n;
    out.nconstraints = C;
    out.words = W;
    out.viol.assign((size_t)n * W, 0);
    out.err.assign((size_t)n * W, 0);
    out.totals.assign(C, 0);
    out.err_totals.assign(C, 0);
    out.errlist.clear();
    auto t0 = std::chrono::steady_clock::now();
    auto rows_of = [&](uint32_t level) -> uint32_t { return level == 0 ? n : b->scope_rows[level]; };
    std::vector<std::vector<uint8_t>> slot(c.slot_level.size());
    for (size_t s = 0; s < slot.size(); ++s) slot[s].assign(rows_of(c.slot_level[s]), 0);
    for (const GkOp& op : c.ops) {
      const uint32_t kind = op.w0 & 0xffu, level = (op.w0 >> 8) & 0xffu, o = op.w0 >> 16;
      if (kind == GK_N_END) break;
      if (kind == GK_N_PHASE) continue;
      switch (kind) {
        case GK_N_CONST:
          std::fill(slot[o].begin(), slot[o].end(), (uint8_t)(op.w1 & 1));
          break;
        case GK_N_ATOM: {
          const GkColumn& col = h.cols[op.w1 >> 8];
          const uint32_t aop = op.w1 & 0xffu, R = rows_of(level);
          for (uint32_t r = 0; r < R; ++r) slot[o][r] = gk_atom(col, r, aop, op.w2, op.w3, c.pool.data(), c.cbytes.data());
          break;
        }
        case GK_N_ATOMS: {
          const GkColumn& col = h.cols[op.w1 >> 8];
          const uint32_t R = rows_of(level);
          for (uint32_t j = 0; j < op.w3; ++j) {
            const uint32_t* e = &c.pool[op.w2 + j * GK_ATOMS_ENT];
            auto& dst = slot[e[0] >> 16];
            for (uint32_t r = 0; r < R; ++r) dst[r] = gk_atom(col, r, e[0] & 0xffu, e[1], e[2], c.pool.data(), c.cbytes.data());
          }
          break;
        }
        case GK_N_GATE: {
          const uint32_t f = op.w2, R = rows_of(level);
          for (uint32_t r = 0; r < R; ++r) {
            bool v = !(f & GK_G_OR);
            for (uint32_t j = 0; j < op.w3; ++j) {
              const uint32_t e = c.pool[op.w1 + j];
              bool x = slot[e & 0xffffu][r] ^ ((e >> 31) != 0);
              v = (f & GK_G_OR) ? (v || x) : (v && x);
            }
            slot[o][r] = v ^ ((f & GK_G_NEG_OUT) != 0);
          }
          break;
        }
        case GK_N_BCAST: {
          const auto& off = b->scope_off[level];
          for (uint32_t j = 0; j < op.w3; ++j) {
            const uint32_t e = c.pool[op.w1 + j];
            const auto& in = slot[e & 0xffffu];
            auto& dst = slot[e >> 16];
            for (size_t p = 0; p + 1 < off.size(); ++p)
              for (uint32_t r = off[p]; r < off[p + 1]; ++r) dst[r] = in[p];
          }
          break;
        }
        case GK_N_ACC: {
          const auto& off = b->scope_off[level];
          for (uint32_t j = 0; j < op.w3; ++j) {
            const uint32_t e = c.pool[op.w1 + j];
            const auto& in = slot[e & 0xffffu];
            auto& dst = slot[e >> 16];
            for (size_t p = 0; p + 1 < off.size(); ++p) {
              bool any = false;
              for (uint32_t r = off[p]; r < off[p + 1]; ++r) any = any || in[r];
              dst[p] = any;
            }
          }
          break;
        }
        case GK_N_MATCH: {
          auto& err = slot[op.w1 & 0xffffu];
          for (uint32_t obj = 0; obj < n; ++obj) {
            int m = (h.flags[obj] & GK_F_SKIP) ? 0 : gk_match(h, c.pool.data(), c.cbytes.data(), c.match[op.w2], obj);
            slot[o][obj] = m > 0;
            err[obj] = m < 0;
            if (m < 0) {
              out.errlist.push_back(obj);
              out.errlist.push_back(op.w2);
              out.errlist.push_back((uint32_t)-m);
            }
          }
          break;
        }
        default: throw BackendError{"hostemu: unknown netlist op"};
      }
    }
    for (uint32_t cix = 0; cix < C; ++cix) {
      if (!active[cix]) continue;
      const GkOutEnt& oe = c.outs[cix];
      const auto &prog = slot[oe.prog_slot], &mt = slot[oe.match_slot], &er = slot[oe.err_slot];
      for (uint32_t obj = 0; obj < n; ++obj) {
        bool pv = (oe.flags & 1) ? true : (oe.flags & 2) ? false : prog[obj] != 0;
        if (pv && mt[obj]) {
          out.viol[(size_t)obj * W + cix / 32] |= 1u << (cix & 31);
          out.totals[cix]++;
        }
        if (er[obj]) {
          out.err[(size_t)obj * W + cix / 32] |= 1u << (cix & 31);
          out.err_totals[cix]++;
        }
      }
    }
    out.kernel_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    out.launches = 0;
  }
  // "device" buffers are host memory here: the same address arithmetic as the CUDA backend (bitmap shard and totals of this
  // rank stored into every peer's receive buffer), so the C ABI's peer addressing can be tested without GPUs
  void eval_into(void* b, const std::vector<uint32_t>& active, const DevOutPtrs& dst) override {
    EvalOut out;
    eval(b, active, out, true);
    const size_t nw = out.viol.size();
    if (dst.err && nw) memcpy(dst.err, out.err.data(), nw * 4);
    if (dst.totals) memcpy(dst.totals, out.totals.data(), out.totals.size() * 8);
    if (dst.err_totals) memcpy(dst.err_totals, out.err_totals.data(), out.err_totals.size() * 8);
    if (dst.npeers == 0) {
      if (dst.viol && nw) memcpy(dst.viol, out.viol.data(), nw * 4);
      return;
    }
    for (uint32_t q = 0; q < dst.npeers; ++q) {
      if (nw) memcpy(reinterpret_cast<void*>(dst.peer_viol[q]), out.viol.data(), nw * 4);
      auto* t = reinterpret_cast<unsigned long long*>(dst.peer_tot[q]);
      for (size_t c = 0; c < out.totals.size(); ++c) {
        t[c] = out.totals[c];
        t[dst.tot_stride + c] = out.err_totals[c];
      }
    }
  }

 private:
  struct Dict {
    std::vector<uint32_t> off;
    std::vector<uint8_t> bytes;
  };
  std::mutex mu_;
  const Compiled* prog_ = nullptr;
  std::shared_ptr<const Dict> dict_;
};

Backend* make_backend(int) { return new HostEmuBackend(); }

}  // namespace gk
