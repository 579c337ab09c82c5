// TEST-ONLY backend: runs the per-object core (gatekeeper_b200/csrc/vm_core.h) in a plain CPU loop so the
// lowering + flattening logic can be unit-tested in the authoring container, which has no GPU.
// It is linked ONLY into tests/_hostemu/libgk_hostemu.so.  The product library
// (gatekeeper_b200/libgk_engine.so) links kernels.cu instead and has no CPU path at all.
#include <algorithm>
#include <chrono>

#include "../../gatekeeper_b200/csrc/backend.hpp"
#include "../../gatekeeper_b200/csrc/vm_core.h"

namespace gk {

struct EmuBatch {
  PackedBatch pb;
  GkBatch hdr{};
  uint32_t n = 0, words = 1;
};

class HostEmuBackend : public Backend {
 public:
  const char* name() const override { return "hostemu-TEST-ONLY"; }
  void set_program(const Compiled& c) override { prog_ = &c; }
  void sync_strings(const StringTable& st) override { st.snapshot(dict_off_, dict_bytes_); }
  void* upload(const HostBatch& hb, const Compiled& c, double* ms, uint64_t* bytes) override {
    auto* b = new EmuBatch();
    pack_batch(hb, c, b->pb);
    b->hdr = rebase_batch(b->pb, b->pb.arena.data(), b->pb.arena.data());
    b->n = hb.n;
    b->words = std::max<uint32_t>(1, (uint32_t)((c.cons.size() + 31) / 32));
    if (ms) *ms = 0;
    if (bytes) *bytes = b->pb.arena.size();
    return b;
  }
  void release(void* b) override { delete static_cast<EmuBatch*>(b); }
  void eval(void* bb, const std::vector<uint32_t>& active, EvalOut& out, bool) override {
    auto* b = static_cast<EmuBatch*>(bb);
    const Compiled& c = *prog_;
    const uint32_t C = (uint32_t)c.cons.size(), W = b->words;
    GkBatch h = b->hdr;
    h.dict_off = dict_off_.data();
    h.dict_bytes = dict_bytes_.data();
    h.dict_n = (uint32_t)dict_off_.size() - 1;
    out.n = b->n;
    out.nconstraints = C;
    out.words = W;
    out.viol.assign((size_t)b->n * W, 0);
    out.err.assign((size_t)b->n * W, 0);
    out.totals.assign(C, 0);
    out.err_totals.assign(C, 0);
    out.errlist.clear();
    auto t0 = std::chrono::steady_clock::now();
    const GkColumn* cols = h.cols;
    const GkScope* scopes = h.scopes;
    for (uint32_t obj = 0; obj < b->n; ++obj) {
      if (h.flags[obj] & GK_F_SKIP) continue;
      unsigned long long cse = 0, cse_valid = 0;
      uint32_t cur_mid = GK_NONE;
      int mres = 0;
      for (uint32_t cix = 0; cix < C; ++cix) {
        if (!active[cix]) continue;
        const GkCons& cc = c.cons[cix];
        if (cc.match_id != cur_mid) {
          cur_mid = cc.match_id;
          mres = gk_match(h, c.pool.data(), c.cbytes.data(), c.match[cur_mid], obj);
        }
        int code = mres < 0 ? -mres : 0;
        int flag = 0;
        bool v = cc.pc == GK_PC_ACCEPT ? true
                 : cc.pc == GK_PC_REJECT ? false
                                         : gk_eval_prog(cols, scopes, c.instr.data(), c.pool.data(), c.cbytes.data(), cc.pc, obj, true, cse, cse_valid, &flag);
        v = v && mres > 0;
        if (mres > 0 && flag) {
          v = false;
          code = flag;
        }
        if (code) {
          out.err[(size_t)obj * W + cix / 32] |= 1u << (cix & 31);
          out.err_totals[cix]++;
          out.errlist.push_back(obj);
          out.errlist.push_back(cix);
          out.errlist.push_back((uint32_t)code);
        }
        if (v) {
          out.viol[(size_t)obj * W + cix / 32] |= 1u << (cix & 31);
          out.totals[cix]++;
        }
      }
    }
    out.kernel_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    out.launches = 0;
  }
  void eval_into(void*, const std::vector<uint32_t>&, const DevOutPtrs&) override {
    throw BackendError{"hostemu has no device buffers"};
  }

 private:
  const Compiled* prog_ = nullptr;
  std::vector<uint32_t> dict_off_;
  std::vector<uint8_t> dict_bytes_;
};

Backend* make_backend(int) { return new HostEmuBackend(); }

}  // namespace gk
