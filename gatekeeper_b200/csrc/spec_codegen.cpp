// Netlist -> CUDA C++ text (see spec_codegen.hpp).  The semantics of every op are those of the interpreter (tile_kernel.cuh) and of
// the test-only host emulation; tests/_hostemu compiles the SAME text with g++ (-DGK_SPEC_HOST) and compares it object by object
// with the interpreted netlist (GK_SPEC_CHECK=1), which is how the generator is verified in a container without a GPU.
#include "spec_codegen.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <set>
#include <sstream>
#include <tuple>

#include "spec_headers.inc"   // kSpecHdrProgram / kSpecHdrVmCore: the text of program.h and vm_core.h (written by build.py)

namespace gk {

namespace {

std::string strip_includes(const char* text) {
  std::istringstream in(text);
  std::string line, out;
  while (std::getline(in, line)) {
    if (line.rfind("#pragma once", 0) == 0 || line.rfind("#include \"program.h\"", 0) == 0) continue;
    out += line;
    out += '\n';
  }
  return out;
}

std::string hx(uint32_t v) {
  char b[24];
  snprintf(b, sizeof b, "0x%xu", v);
  return b;
}

// the fixed part: launch wrapper (device) / per-object entry point (host test build)
const char* kWrapper = R"GKSRC(
#ifndef GK_SPEC_HOST
__device__ __forceinline__ uint32_t gk_tr_step(uint32_t x, uint32_t lane, uint32_t j, uint32_t m) {
  const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
  return (lane & j) ? ((x & ~m) | ((y & ~m) >> j)) : ((x & m) | ((y & m) << j));
}
// 32 x 32 bit transpose across the warp: lane i gives row i and receives column i
__device__ __forceinline__ uint32_t gk_tr32(uint32_t x, uint32_t lane) {
  x = gk_tr_step(x, lane, 16u, 0x0000ffffu);
  x = gk_tr_step(x, lane, 8u, 0x00ff00ffu);
  x = gk_tr_step(x, lane, 4u, 0x0f0f0f0fu);
  x = gk_tr_step(x, lane, 2u, 0x33333333u);
  x = gk_tr_step(x, lane, 1u, 0x55555555u);
  return x;
}

extern "C" __global__ void __launch_bounds__(GK_SPEC_THREADS, GK_SPEC_MINB) gk_spec_kernel(const __grid_constant__ GkKParams p) {
  extern __shared__ uint4 gk_smem4[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(gk_smem4);
  const uint32_t NC = p.batch.ncols, NS = p.batch.nscopes;
  GkColumn* cols = reinterpret_cast<GkColumn*>(smem);
  GkScope* scopes = reinterpret_cast<GkScope*>(smem + (size_t)NC * sizeof(GkColumn));
  uint32_t* act = reinterpret_cast<uint32_t*>(smem + (size_t)NC * sizeof(GkColumn) + (size_t)NS * sizeof(GkScope));
  uint32_t* s_tot = act + GK_SPEC_W * 32u;
  uint32_t* s_err = s_tot + GK_SPEC_W * 32u;
  uint32_t* actw = s_err + GK_SPEC_W * 32u;   // the enforcement-point mask packed: bit c of word c / 32
  {
    const uint4* a = reinterpret_cast<const uint4*>(p.batch.cols);
    uint4* d = reinterpret_cast<uint4*>(cols);
    for (uint32_t i = threadIdx.x; i < NC * (uint32_t)(sizeof(GkColumn) / 16); i += blockDim.x) d[i] = a[i];
    const uint4* b = reinterpret_cast<const uint4*>(p.batch.scopes);
    uint4* e = reinterpret_cast<uint4*>(scopes);
    for (uint32_t i = threadIdx.x; i < NS * (uint32_t)(sizeof(GkScope) / 16); i += blockDim.x) e[i] = b[i];
    for (uint32_t i = threadIdx.x; i < GK_SPEC_W * 32u; i += blockDim.x) {
      act[i] = i < GK_SPEC_C ? p.active[i] : 0u;
      s_tot[i] = 0u;
      s_err[i] = 0u;
    }
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t i = threadIdx.x; i < GK_SPEC_W * 32u; i += blockDim.x) {   // (whole warps: both bounds are multiples of 32)
    const uint32_t bits = __ballot_sync(0xffffffffu, act[i] != 0u);
    if (lane == 0u) actw[i >> 5] = bits;
  }
  __syncthreads();
  uint32_t tv[GK_SPEC_W], te[GK_SPEC_W];
#pragma unroll
  for (uint32_t w = 0; w < GK_SPEC_W; ++w) tv[w] = te[w] = 0u;
  for (uint32_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
    const uint32_t obj0 = t * p.tile, nobj = min(p.tile, p.batch.n - obj0);
    // An object with more rows in a scope than a mask register holds makes gk_spec_object() return true: the tile is then listed
    // for the interpreter, which evaluates it again and overwrites what this kernel stored for it; its totals are not counted here.
    int big = 0;
    uint32_t tvt[GK_SPEC_W], tet[GK_SPEC_W];
#pragma unroll
    for (uint32_t w = 0; w < GK_SPEC_W; ++w) tvt[w] = tet[w] = 0u;
    for (uint32_t base = 0; base < nobj; base += blockDim.x) {
      const uint32_t o = base + threadIdx.x;
      uint32_t vw[GK_SPEC_W], ew[GK_SPEC_W];
#pragma unroll
      for (uint32_t w = 0; w < GK_SPEC_W; ++w) vw[w] = ew[w] = 0u;
      if (o < nobj) {
        if (gk_spec_object(p.batch, cols, scopes, p.prog.pool, p.prog.cbytes, actw, p.out, obj0 + o, vw, ew)) big = 1;
        const size_t at = (size_t)(obj0 + o) * GK_SPEC_W;
#if GK_SPEC_W == 2
        const uint2 v2 = make_uint2(vw[0], vw[1]);
        if (p.npeers) {
          for (uint32_t q = 0; q < p.npeers; ++q) *reinterpret_cast<uint2*>(p.peer_viol[q] + at) = v2;   // this rank is one of the peers
        } else {
          *reinterpret_cast<uint2*>(p.out.viol + at) = v2;
        }
        *reinterpret_cast<uint2*>(p.out.err + at) = make_uint2(ew[0], ew[1]);
#else
#pragma unroll
        for (uint32_t w = 0; w < GK_SPEC_W; ++w) {
          if (p.npeers) {
            for (uint32_t q = 0; q < p.npeers; ++q) p.peer_viol[q][at + w] = vw[w];
          } else {
            p.out.viol[at + w] = vw[w];
          }
          p.out.err[at + w] = ew[w];
        }
#endif
      }
      // totals: after the transpose lane c holds the bit of constraint (32 w + c) for the warp's 32 objects
      uint32_t anye = 0u;
#pragma unroll
      for (uint32_t w = 0; w < GK_SPEC_W; ++w) {
        tvt[w] += (uint32_t)__popc(gk_tr32(vw[w], lane));
        anye |= ew[w];
      }
      if (__any_sync(0xffffffffu, anye != 0u)) {
#pragma unroll
        for (uint32_t w = 0; w < GK_SPEC_W; ++w) tet[w] += (uint32_t)__popc(gk_tr32(ew[w], lane));
      }
    }
    if (__syncthreads_or(big)) {
      if (threadIdx.x == 0) p.tile_list[atomicAdd(p.tile_count, 1u)] = t;
    } else {
#pragma unroll
      for (uint32_t w = 0; w < GK_SPEC_W; ++w) {
        tv[w] += tvt[w];
        te[w] += tet[w];
      }
    }
  }
#pragma unroll
  for (uint32_t w = 0; w < GK_SPEC_W; ++w) {
    if (tv[w]) atomicAdd(&s_tot[w * 32u + lane], tv[w]);
    if (te[w]) atomicAdd(&s_err[w * 32u + lane], te[w]);
  }
  __syncthreads();
  for (uint32_t c = threadIdx.x; c < GK_SPEC_C; c += blockDim.x) {
    if (s_tot[c]) atomicAdd(p.out.totals + c, (unsigned long long)s_tot[c]);
    if (s_err[c]) atomicAdd(p.out.err_totals + c, (unsigned long long)s_err[c]);
  }
}
#else
// TEST-ONLY host build (tests/_hostemu, GK_SPEC_CHECK=1): one object; returns 1 when the object does not fit the mask registers
extern "C" int gk_spec_host(const GkKParams* p, uint32_t obj, uint32_t* vw, uint32_t* ew) {
  uint32_t actw[GK_SPEC_W];
  for (uint32_t w = 0; w < GK_SPEC_W; ++w) vw[w] = ew[w] = actw[w] = 0u;
  for (uint32_t c = 0; c < GK_SPEC_C; ++c)
    if (p->active[c]) actw[c >> 5] |= 1u << (c & 31u);
  return gk_spec_object(p->batch, p->batch.cols, p->batch.scopes, p->prog.pool, p->prog.cbytes, actw, p->out, obj, vw, ew) ? 1 : 0;
}
#endif
)GKSRC";

// ---- one spec.match block written out as gk_spec_match_<mid>(): the same eight criteria in the same order as gk_match_row() /
// gk_match() (vm_core.h; reference pkg/mutation/match/match.go:32-65, pkg/target/matcher.go:21-71), with everything the block fixes
// folded in.  Returns false (nothing written) when the block's tables do not lie inside the pool: the caller then keeps the generic form.
// a wildcard pattern (wildcard.go:17-41) against the bytes s[0, sl): the pattern's bytes are immediates
std::string wild_expr(const Compiled& c, uint32_t mode, uint32_t off, uint32_t len, const std::string& s, const std::string& sl, bool generate_name) {
  const std::string L = std::to_string(len) + "u";
  if (generate_name && mode != GK_W_PREFIX && mode != GK_W_CONTAINS) return "false";   // exact and "*x" never match a generateName
  const std::string generic = std::string(generate_name ? "gk_wild_gen(" : "gk_wild(") + std::to_string(mode) + "u, cbytes + " + std::to_string(off) + "u, " + L + ", " + s + ", " + sl + ")";
  if ((size_t)off + len > c.cbytes.size() || len > 24u) return generic;
  if (mode == GK_W_CONTAINS) return len == 0 ? "true" : generic;
  std::string e = "(" + sl + (mode == GK_W_EXACT ? " == " : " >= ") + L;
  for (uint32_t i = 0; i < len; ++i) {
    const std::string at = mode == GK_W_SUFFIX ? s + " + (" + sl + " - " + L + ") + " + std::to_string(i) + "u" : s + " + " + std::to_string(i) + "u";
    e += " && GK_LD(" + at + ") == " + hx(c.cbytes[off + i]);
  }
  return e + ")";
}

// labels.Selector.Matches over the (key, value) run [lo, hi) of `kv`: one walk of the run fetches the value of every key the
// selector names (backwards, so that the first entry of a key is the one that stays, as in gk_label()); statements that `return 0`
bool selector_code(std::ostringstream& o, const Compiled& c, uint32_t off, uint32_t nreq, const std::string& kv, const std::string& lo, const std::string& hi,
                   const std::string& tag) {
  struct Req {
    uint32_t key, op;
    std::vector<uint32_t> vals;
  };
  std::vector<Req> reqs;
  std::vector<uint32_t> keys;
  for (uint32_t r = 0; r < nreq; ++r) {
    if ((size_t)off + 3 > c.pool.size()) return false;
    Req q{c.pool[off], c.pool[off + 1], {}};
    const uint32_t nv = c.pool[off + 2];
    if ((size_t)off + 3 + nv > c.pool.size() || nv > 64u) return false;
    q.vals.assign(c.pool.begin() + off + 3, c.pool.begin() + off + 3 + nv);
    off += 3 + nv;
    if (std::find(keys.begin(), keys.end(), q.key) == keys.end()) keys.push_back(q.key);
    reqs.push_back(std::move(q));
  }
  if (reqs.empty()) return true;
  o << "  {\n    uint32_t";
  for (size_t k = 0; k < keys.size(); ++k) o << (k ? ", " : " ") << tag << k << " = GK_NONE";
  o << ";\n    _Pragma(\"unroll 1\") for (uint32_t i = " << hi << "; i-- > " << lo << ";) {\n      const uint32_t k = GK_LD(" << kv << " + 2u * i), v = GK_LD(" << kv
    << " + 2u * i + 1u);\n";
  for (size_t k = 0; k < keys.size(); ++k) o << "      if (k == " << hx(keys[k]) << ") " << tag << k << " = v;\n";
  o << "    }\n";
  for (const Req& q : reqs) {
    const std::string x = tag + std::to_string(std::find(keys.begin(), keys.end(), q.key) - keys.begin());
    const std::string has = "(" + x + " != GK_NONE)";
    std::string in = "false";
    if (!q.vals.empty()) {
      in = "(" + has + " && (";
      for (size_t j = 0; j < q.vals.size(); ++j) in += (j ? " || " : "") + x + " == " + hx(q.vals[j]);
      in += "))";
    }
    const std::string ok = q.op == GK_SEL_IN ? in : q.op == GK_SEL_NOTIN ? "!" + in : q.op == GK_SEL_EXISTS ? has : "!" + has;
    o << "    if (!(" << ok << ")) return 0;\n";
  }
  o << "  }\n";
  return true;
}

bool emit_match(std::ostringstream& out, const Compiled& c, uint32_t mid) {
  const GkMatch& m = c.match[mid];
  const std::string M = std::to_string(mid);
  const char* sig_tail = "(const GkBatch& B, const uint32_t* __restrict__ pool, const uint8_t* __restrict__ cbytes, ";
  std::ostringstream o;
  if (!(m.flags & GK_M_HAS_MATCH)) {   // matcher.go:22-25: no spec.match matches everything
    out << "GK_SPEC_FN int gk_spec_match_" << M << sig_tail << "const uint32_t obj) {\n  (void)B; (void)pool; (void)cbytes; (void)obj;\n  return 1;\n}\n";
    return true;
  }
  o << "GK_SPEC_FN int gk_spec_mrow_" << M << sig_tail << "const uint32_t row, const uint32_t obj) {\n  (void)pool; (void)cbytes; (void)obj;\n";
  o << "  const uint32_t fl = GK_LD(B.flags + row);\n  const bool is_ns = (fl & GK_F_IS_NS) != 0u;\n  (void)is_ns;\n";
  // 1 kinds -- match.go:181-201
  if (m.kinds_n) {
    std::string any;
    uint32_t off = m.kinds_off;
    bool always = false;
    for (uint32_t e = 0; e < m.kinds_n; ++e) {
      if ((size_t)off + 3 > c.pool.size()) return false;
      const uint32_t nk = c.pool[off], ng = c.pool[off + 1], wild = c.pool[off + 2];
      if ((size_t)off + 3 + nk + ng > c.pool.size() || nk > 64u || ng > 64u) return false;
      std::string km, gm;
      if (!(nk == 0 || (wild & 1u))) {
        for (uint32_t j = 0; j < nk; ++j) km += (j ? " || kind == " : "kind == ") + hx(c.pool[off + 3 + j]);
        km = "(" + km + ")";
      }
      if (!(ng == 0 || (wild & 2u))) {
        for (uint32_t j = 0; j < ng; ++j) gm += (j ? " || group == " : "group == ") + hx(c.pool[off + 3 + nk + j]);
        gm = "(" + gm + ")";
      }
      const std::string both = km.empty() && gm.empty() ? "" : km.empty() ? gm : gm.empty() ? km : "(" + km + " && " + gm + ")";
      if (both.empty()) always = true;
      else any += (any.empty() ? "" : " || ") + both;
      off += 3 + nk + ng;
    }
    if (!always) o << "  {\n    const uint32_t kind = GK_LD(B.kind_sid + row), group = GK_LD(B.group_sid + row);\n    (void)kind; (void)group;\n    if (!(" << any << ")) return 0;\n  }\n";
  }
  // 2 scope -- match.go:214-227
  if (m.flags & (GK_M_SCOPE_CLUSTER | GK_M_SCOPE_NAMESPACED)) {
    o << "  {\n    const bool has_ns = (fl & (GK_F_HAS_NS | GK_F_NS_OBJ)) != 0u;\n";
    if (m.flags & GK_M_SCOPE_CLUSTER) o << "    if (!(is_ns || !has_ns)) return 0;\n";
    if (m.flags & GK_M_SCOPE_NAMESPACED) o << "    if (!(!is_ns && has_ns)) return 0;\n";
    o << "  }\n";
  }
  // 3/4 namespaces, excludedNamespaces -- match.go:118-179
  if (m.ns_n || m.exns_n) {
    if ((size_t)m.ns_off + 3 * (size_t)m.ns_n > c.pool.size() || (size_t)m.exns_off + 3 * (size_t)m.exns_n > c.pool.size() || m.ns_n > 64u || m.exns_n > 64u) return false;
    o << "  if (fl & GK_F_NSNAME) {\n    const uint32_t s0 = GK_LD(B.nsn_off + row), sl = GK_LD(B.nsn_off + row + 1u) - s0;\n    const uint8_t* s = B.nsn_bytes + s0;\n    (void)s; (void)sl;\n";
    if (m.ns_n) {
      o << "    if (!(";
      for (uint32_t j = 0; j < m.ns_n; ++j) {
        const uint32_t* e = &c.pool[m.ns_off + 3 * j];
        o << (j ? "\n          || " : "") << wild_expr(c, e[0], e[1], e[2], "s", "sl", false);
      }
      o << ")) return 0;\n";
    }
    for (uint32_t j = 0; j < m.exns_n; ++j) {
      const uint32_t* e = &c.pool[m.exns_off + 3 * j];
      o << "    if (" << wild_expr(c, e[0], e[1], e[2], "s", "sl", false) << ") return 0;\n";
    }
    o << "  }\n";
  }
  // 5 labelSelector -- match.go:103-116
  if (m.flags & GK_M_HAS_LSEL) {
    if (m.flags & GK_M_LSEL_INVALID) {
      o << "  return -GK_E_LSEL_INVALID;\n}\n";
      goto wrapper;
    }
    o << "  {\n    const uint32_t l0 = GK_LD(B.lbl_off + row), l1 = GK_LD(B.lbl_off + row + 1u);\n";
    if (!selector_code(o, c, m.lsel_off, m.lsel_n, "B.lbl_kv", "l0", "l1", "lv")) return false;
    o << "  }\n";
  }
  // 6 namespaceSelector -- match.go:73-101
  if (m.flags & GK_M_HAS_NSSEL) {
    o << "  {\n    const bool ns_obj = (fl & GK_F_NS_OBJ) != 0u, obj_ns = (fl & GK_F_HAS_NS) != 0u;\n    if (is_ns || ns_obj || obj_ns) {\n";
    if (m.flags & GK_M_NSSEL_INVALID) {
      o << "      return -GK_E_NSSEL_INVALID;\n    }\n  }\n";
    } else {
      o << "      if (is_ns) {\n        const uint32_t l0 = GK_LD(B.lbl_off + row), l1 = GK_LD(B.lbl_off + row + 1u);\n";
      if (!selector_code(o, c, m.nssel_off, m.nssel_n, "B.lbl_kv", "l0", "l1", "sv")) return false;
      o << "      } else {\n        if (!ns_obj) return -GK_E_NS_MISSING;\n        const uint32_t nr = GK_LD(B.nsrow + obj);\n        const uint32_t l0 = GK_LD(B.nsl_off + nr), l1 = GK_LD(B.nsl_off + nr + 1u);\n";
      if (!selector_code(o, c, m.nssel_off, m.nssel_n, "B.nsl_kv", "l0", "l1", "nv")) return false;
      o << "      }\n    }\n  }\n";
    }
  }
  // 7 name -- match.go:203-212
  if (m.flags & GK_M_HAS_NAME) {
    o << "  {\n    const uint32_t a = GK_LD(B.name_off + row), al = GK_LD(B.name_off + row + 1u) - a;\n    const uint8_t* nm = B.name_bytes + a;\n    (void)nm; (void)al;\n";
    o << "    bool ok = " << wild_expr(c, m.name_mode, m.name_boff, m.name_len, "nm", "al", false) << ";\n";
    const std::string ge = wild_expr(c, m.name_mode, m.name_boff, m.name_len, "gn", "gl", true);
    if (ge != "false")
      o << "    if (!ok) {\n      const uint32_t g = GK_LD(B.gen_off + row), gl = GK_LD(B.gen_off + row + 1u) - g;\n      const uint8_t* gn = B.gen_bytes + g;\n      (void)gn; (void)gl;\n      ok = " << ge
        << ";\n    }\n";
    o << "    if (!ok) return 0;\n  }\n";
  }
  // 8 source -- match.go:229-253
  {
    if (m.flags & GK_M_SRC_INVALID) {
      o << "  return -GK_E_SRC_INVALID_MATCH;\n}\n";
      goto wrapper;
    }
    const uint32_t msrc = (m.flags >> GK_M_SRC_SHIFT) & 7u;
    if (msrc != GK_SRC_ALL) {
      o << "  {\n    const uint32_t tsrc = (fl & GK_F_SRC_MASK) >> GK_F_SRC_SHIFT;\n    if (tsrc == GK_SRC_EMPTY) return -GK_E_SRC_UNSPECIFIED;\n    if (tsrc == GK_SRC_INVALID) return -GK_E_SRC_INVALID_OBJ;\n    if (tsrc != "
        << msrc << "u) return 0;\n  }\n";
    }
  }
  o << "  return 1;\n}\n";
wrapper:
  // Matcher.Match: the object, then the old object (matcher.go:21-71); one copy of the row code for both
  o << "GK_SPEC_FN int gk_spec_match_" << M << sig_tail << "const uint32_t obj) {\n  int nil = 0;\n"
    << "  _Pragma(\"unroll 1\") for (uint32_t pass = 0u; pass < 2u; ++pass) {\n"
    << "    if (pass && !B.has_old) {\n      ++nil;\n      break;\n    }\n"
    << "    const uint32_t row = pass ? B.n + obj : obj;\n"
    << "    if (!(GK_LD(B.flags + row) & GK_F_HAS_OBJ)) {\n      ++nil;\n      continue;\n    }\n"
    << "    const int r = gk_spec_mrow_" << M << "(B, pool, cbytes, row, obj);\n"
    << "    if (r < 0) return pass ? r - GK_E_FROM_OLD : r;   // the error text names the object that failed: here the old one\n"
    << "    if (r) return r;\n  }\n  return nil == 2 ? -GK_E_NO_OBJECT : 0;\n}\n";
  out << o.str();
  return true;
}

// One netlist op after SSA renaming (slots are reused by liveness; every write gets a fresh variable): its text, what it
// reads and what it defines.  Nodes are emitted depth-first from the constraint results (see spec_codegen()), not in netlist
// order, which keeps the live values of one scope subtree together instead of all ~300 at once.
struct SpecNode {
  std::string code;
  std::vector<uint32_t> deps;
  bool done = false;
};

struct Gen {
  const Compiled& c;
  std::vector<SpecNode> nodes;
  std::vector<int> producer;               // variable -> node, or -(scope + 1) for an atom of that scope's row loop
  struct AtomRec {
    uint32_t var, scope;
    std::string expr;                              // boolean expression over the row's loaded values
    std::map<uint32_t, uint32_t> need;             // column -> encodings the expression reads
    int group = -1;                                // row loop it is computed in (see spec_codegen())
  };
  std::vector<AtomRec> atoms;
  std::vector<int> atom_of;                        // variable -> atom record, or -1
  std::map<uint32_t, uint32_t>* need_now = nullptr;
  std::vector<uint8_t> scope_used;
  std::vector<int> cur;                    // slot -> variable holding its current value (-1: never written)
  uint32_t nvar = 0;
  size_t n_fast = 0, n_generic = 0;
  std::set<uint32_t> match_used;
  std::ostringstream body;                 // the node being built
  std::vector<uint32_t> deps;

  explicit Gen(const Compiled& cc) : c(cc) {
    const size_t NS = c.schema.scopes.size();
    scope_used.assign(NS, 0);
    cur.assign(c.slot_level.size() + 1, -1);
  }
  void use_scope(uint32_t s) {
    while (!scope_used[s]) {
      scope_used[s] = 1;
      if (s == 0) break;
      s = (uint32_t)c.schema.scopes[s].parent;
    }
  }
  std::string rd(uint32_t slot) {
    if (slot >= cur.size() || cur[slot] < 0) return "0u";
    deps.push_back((uint32_t)cur[slot]);
    return "v" + std::to_string(cur[slot]);
  }
  uint32_t fresh(uint32_t slot, int prod) {
    if (slot >= cur.size()) cur.resize(slot + 1, -1);
    cur[slot] = (int)nvar;
    producer.push_back(prod);
    atom_of.push_back(-1);
    return nvar++;
  }
  uint32_t fresh(uint32_t slot) { return fresh(slot, (int)nodes.size()); }   // defined by the node being built
  void finish_node() {
    SpecNode n;
    n.code = body.str();
    n.deps = deps;
    nodes.push_back(std::move(n));
    body.str("");
    deps.clear();
  }

  // a boolean C expression for one atom on `row` of column ci; the loads it needs are recorded in col_need
  std::string atom_expr(uint32_t level, uint32_t ci, uint32_t aop, uint32_t a, uint32_t b) {
    const std::string K = std::to_string(ci);
    (void)level;
    auto need = [&](uint32_t enc) { (*need_now)[ci] |= enc; };
    const std::string t = "t" + K, i = "i" + K, n = "n" + K;
    switch (aop) {
      case GK_OP_TRUTHY: need(GK_ENC_VT); ++n_fast; return "(" + t + " != 0u && " + t + " != 2u)";
      case GK_OP_DEFINED: need(GK_ENC_VT); ++n_fast; return "(" + t + " != 0u)";
      case GK_OP_VTMASK: need(GK_ENC_VT); ++n_fast; return "(((" + hx(a) + " >> " + t + ") & 1u) != 0u)";
      case GK_OP_SID_EQ: need(GK_ENC_SID); ++n_fast; return "(" + i + " == " + hx(a) + ")";
      case GK_OP_SID_IN:
        if (b <= 16u && (size_t)a + b <= c.pool.size()) {
          need(GK_ENC_SID);
          ++n_fast;
          if (b == 0) return "false";
          std::string e = "(";
          for (uint32_t j = 0; j < b; ++j) e += (j ? " || " : "") + i + " == " + hx(c.pool[a + j]);
          return e + ")";
        }
        break;
      case GK_OP_NUM_CMP: {
        if ((size_t)a + 1 >= c.pool.size()) break;
        const int64_t k = (int64_t)(((uint64_t)c.pool[a + 1] << 32) | c.pool[a]);
        if (k == INT64_MIN || k == INT64_MAX) break;   // the sentinels of non-numbers: the generic path orders by type rank
        // (a defined non-number holds INT64_MIN / INT64_MAX by its type rank, so one signed compare is OPA's order: tile_kernel.cuh)
        need(GK_ENC_VT | GK_ENC_NUM);
        ++n_fast;
        const char* opn = b == GK_CMP_LT ? "<" : b == GK_CMP_LE ? "<=" : b == GK_CMP_GT ? ">" : b == GK_CMP_GE ? ">=" : b == GK_CMP_EQ ? "==" : "!=";
        return "(" + t + " != 0u && " + n + " " + opn + " " + std::to_string((long long)k) + "LL)";
      }
      case GK_OP_ANYPREFIX: {
        bool all_short = (size_t)a + (size_t)b * GK_PREFIX_ENT <= c.pool.size();
        for (uint32_t j = 0; all_short && j < b; ++j) all_short = c.pool[a + j * GK_PREFIX_ENT] <= GK_HEAD_BYTES;
        if (!all_short) break;
        need(GK_ENC_VT | GK_ENC_HEAD);
        ++n_fast;
        if (b == 0) return "false";
        const std::string h = "h" + K;
        static const char* comp[8] = {"a.x", "a.y", "a.z", "a.w", "b.x", "b.y", "b.z", "b.w"};
        std::string e = "(" + t + " == 5u && (";
        for (uint32_t j = 0; j < b; ++j) {
          const uint32_t* ent = &c.pool[a + j * GK_PREFIX_ENT];
          std::string one = "((" + h + "b.w >> 24) >= " + std::to_string(ent[0]) + "u";
          for (uint32_t w = 0; w < GK_HEAD_WORDS; ++w) {
            const uint32_t m = ent[2 + GK_HEAD_WORDS + w], v = ent[2 + w] & m;
            if (!m) continue;
            if (m == 0xffffffffu) one += " && " + h + comp[w] + " == " + hx(v);
            else one += " && (" + h + comp[w] + " & " + hx(m) + ") == " + hx(v);
          }
          e += (j ? " || " : "") + one + ")";
        }
        return e + "))";
      }
      default: break;
    }
    ++n_generic;
    return "gk_atom(cols[" + K + "], (uint32_t)row, " + std::to_string(aop) + "u, " + hx(a) + ", " + hx(b) + ", pool, cbytes)";
  }

  void atom(uint32_t level, uint32_t out_slot, uint32_t ci, uint32_t aop, uint32_t a, uint32_t b) {
    use_scope(level);
    AtomRec r;
    r.var = fresh(out_slot, -1);
    r.scope = level;
    need_now = &r.need;
    r.expr = atom_expr(level, ci, aop, a, b);
    need_now = nullptr;
    atom_of[r.var] = (int)atoms.size();
    atoms.push_back(std::move(r));
  }

  // `rm` = the mask of the children of parent row pj inside the object's rows of scope L, `pb` = the parent row's own bit.  The CSR
  // offsets are monotone, so rm = (rows below the END of pj) & ~(rows below its START), and the start of pj is the end of pj - 1: one
  // load and one mask per turn, carried.  (lo<L> is by definition the offset at lo<P>: the first parent's children start at bit 0.)
  // Not unrolled: parents have one to three rows here, the unrolled-by-four body the compiler makes is text that never runs.
  std::string range_loop_head(uint32_t L) {
    const uint32_t P = (uint32_t)c.schema.scopes[L].parent;
    const std::string l = std::to_string(L), p = std::to_string(P);
    return "  {\n  uint32_t lmp = 0u;\n  _Pragma(\"unroll 1\") for (uint32_t pj = 0; pj < n" + p + "; ++pj) {\n    const uint32_t rb = GK_SPEC_LD(o" + l + " + lo" + p + " + pj + 1u) - lo" + l +
           ";\n    const uint32_t lm = rb >= 32u ? 0xffffffffu : ((1u << rb) - 1u), rm = lm & ~lmp, pb = 1u << pj;\n    lmp = lm;\n    (void)pb;\n";
  }

  void op(const GkOp& op) {
    const uint32_t kind = op.w0 & 0xffu, level = (op.w0 >> 8) & 0xffu, out = op.w0 >> 16;
    const std::string L = std::to_string(level);
    switch (kind) {
      case GK_N_ATOM: atom(level, out, op.w1 >> 8, op.w1 & 0xffu, op.w2, op.w3); break;
      case GK_N_ATOMS:
        for (uint32_t j = 0; j < op.w3; ++j) {
          const uint32_t* e = &c.pool[op.w2 + j * GK_ATOMS_ENT];
          atom(level, e[0] >> 16, op.w1 >> 8, e[0] & 0xffu, e[1], e[2]);
        }
        break;
      case GK_N_CONST: {
        use_scope(level);
        const uint32_t v = fresh(out);
        body << "  const uint32_t v" << v << " = " << ((op.w1 & 1u) ? "f" + L : std::string("0u")) << ";\n";
        break;
      }
      case GK_N_GATE: {
        use_scope(level);
        const bool is_or = (op.w2 & GK_G_OR) != 0u, neg_out = (op.w2 & GK_G_NEG_OUT) != 0u;
        std::string e;
        bool any_neg = neg_out;
        for (uint32_t j = 0; j < op.w3; ++j) {
          const uint32_t in = c.pool[op.w1 + j];
          const bool neg = (in >> 31) != 0u;
          any_neg = any_neg || neg;
          e += (j ? (is_or ? " | " : " & ") : "") + std::string(neg ? "~" : "") + rd(in & 0xffffu);
        }
        if (op.w3 == 0) e = is_or ? "0u" : "0xffffffffu", any_neg = true;
        if (neg_out) e = "~(" + e + ")";
        if (any_neg) e = "(" + e + ") & f" + L;
        const uint32_t v = fresh(out);
        body << "  const uint32_t v" << v << " = " << e << ";\n";
        break;
      }
      case GK_N_BCAST: {   // parent-level values -> the rows of the child scope `level`
        use_scope(level);
        const uint32_t P = (uint32_t)c.schema.scopes[level].parent;
        std::vector<std::pair<std::string, uint32_t>> pr;
        for (uint32_t j = 0; j < op.w3; ++j) {
          const uint32_t e = c.pool[op.w1 + j];
          const std::string in = rd(e & 0xffffu);
          pr.emplace_back(in, fresh(e >> 16));
        }
        if (P == 0) {
          for (auto& q : pr) body << "  const uint32_t v" << q.second << " = (0u - (" << q.first << " & 1u)) & f" << L << ";\n";
        } else {
          for (auto& q : pr) body << "  uint32_t v" << q.second << " = 0u;\n";
          body << range_loop_head(level);
          for (auto& q : pr) body << "    if (" << q.first << " & pb) v" << q.second << " |= rm;\n";
          body << "  }\n  }\n";
        }
        break;
      }
      case GK_N_ACC:
      case GK_N_ACC2: {    // EXISTS (at least one / at least two rows of the child scope `level`) per parent row
        use_scope(level);
        const uint32_t P = (uint32_t)c.schema.scopes[level].parent;
        const bool two = kind == GK_N_ACC2;
        std::vector<std::pair<std::string, uint32_t>> pr;
        for (uint32_t j = 0; j < op.w3; ++j) {
          const uint32_t e = c.pool[op.w1 + j];
          const std::string in = rd(e & 0xffffu);
          pr.emplace_back(in, fresh(e >> 16));
        }
        auto test = [&](const std::string& x) { return two ? "(GK_SPEC_POPC(" + x + ") >= 2)" : "((" + x + ") != 0u)"; };
        if (P == 0) {
          for (auto& q : pr) body << "  const uint32_t v" << q.second << " = " << test(q.first) << " ? 1u : 0u;\n";
        } else {
          for (auto& q : pr) body << "  uint32_t v" << q.second << " = 0u;\n";
          body << range_loop_head(level);
          for (auto& q : pr) body << "    if (" << test(q.first + " & rm") << ") v" << q.second << " |= pb;\n";
          body << "  }\n  }\n";
        }
        break;
      }
      case GK_N_MATCH: {
        const uint32_t mid = op.w2;
        match_used.insert(mid);
        const uint32_t vm = fresh(out), ve = fresh(op.w1 & 0xffffu);
        body << "  const int m" << vm << " = skip ? 0 : GK_SPEC_MATCH(" << mid << ", B, pool, cbytes, obj);\n";
        body << "  if (m" << vm << " < 0) GK_SPEC_ERR(obj, " << mid << "u, (uint32_t)(-m" << vm << "));\n";
        body << "  const uint32_t v" << vm << " = m" << vm << " > 0 ? 1u : 0u, v" << ve << " = m" << vm << " < 0 ? 1u : 0u;\n";
        break;
      }
      default: break;
    }
    if (kind != GK_N_ATOM && kind != GK_N_ATOMS) finish_node();
  }
};

}  // namespace

SpecSource spec_codegen(const Compiled& c) {
  SpecSource out;
  const uint32_t C = (uint32_t)c.cons_match.size(), W = std::max<uint32_t>(1, (C + 31) / 32);
  const size_t NS = c.schema.scopes.size();
  Gen g(c);
  g.use_scope(0);
  for (const GkOp& op : c.ops) {
    const uint32_t kind = op.w0 & 0xffu;
    if (kind == GK_N_END) break;
    g.op(op);
  }
  std::ostringstream o;
  o << "// generated by spec_codegen.cpp for constraint-set version " << c.version << ": " << C << " constraints, " << c.ops.size() << " netlist ops\n";
  o << "#define GK_SPEC_C " << C << "u\n#define GK_SPEC_W " << W << "\n";
  o << "#ifndef GK_SPEC_THREADS\n#define GK_SPEC_THREADS 512\n#endif\n#ifndef GK_SPEC_MINB\n#define GK_SPEC_MINB 1\n#endif\n";
  o << "#ifndef GK_SPEC_HOST\n#define GK_LD(p) __ldg(p)   /* pool, cbytes and every batch array are global and read-only here */\n#endif\n";
  o << strip_includes(kSpecHdrProgram) << strip_includes(kSpecHdrVmCore);
  o << R"GKSRC(
#ifdef GK_SPEC_X_NOMATCH   /* (measurement only: what the spec.match pre-filter costs) */
#define GK_SPEC_MATCH(MID, B, pool, cbytes, obj) 1
#else
#define GK_SPEC_MATCH(MID, B, pool, cbytes, obj) gk_spec_match_##MID(B, pool, cbytes, obj)
#endif
#ifdef GK_SPEC_HOST
#define GK_SPEC_FN static inline
#define GK_SPEC_LD(p) (*(p))
#define GK_SPEC_POPC(x) __builtin_popcount(x)
#define GK_SPEC_ERR(o, mid, code)                                    \
  {                                                                  \
    const uint32_t sl_ = (*out.errcount)++;                          \
    if (sl_ < out.errcap) {                                          \
      out.errlist[3 * sl_] = (o);                                    \
      out.errlist[3 * sl_ + 1] = (mid);                              \
      out.errlist[3 * sl_ + 2] = (code);                             \
    }                                                                \
  }
struct uint4 { uint32_t x, y, z, w; };
#else
#define GK_SPEC_FN __device__ __forceinline__
#define GK_SPEC_LD(p) __ldg(p)      /* every array of a resident batch is read-only while it is evaluated */
#define GK_SPEC_POPC(x) __popc(x)
#define GK_SPEC_ERR(o, mid, code)                                    \
  {                                                                  \
    const uint32_t sl_ = atomicAdd(out.errcount, 1u);                \
    if (sl_ < out.errcap) {                                          \
      out.errlist[3 * sl_] = (o);                                    \
      out.errlist[3 * sl_ + 1] = (mid);                              \
      out.errlist[3 * sl_ + 2] = (code);                             \
    }                                                                \
  }
#endif
)GKSRC";
  // ---- the spec.match pre-filter of every block the netlist uses, written out (emit_match below): criteria the block does not have are
  // not emitted, kind / group / label ids and the bytes of the wildcard patterns are immediates, the labels are walked once per selector.
  // GK_SPEC_MATCHGEN=0: the block as a GkMatch literal through the shared gk_match() (vm_core.h) instead -- the form the first
  // generated kernels had; the pool / cbytes reads of that form are what the written-out one removes.
  const bool matchgen = !(getenv("GK_SPEC_MATCHGEN") && atoi(getenv("GK_SPEC_MATCHGEN")) == 0);
  for (uint32_t mid : g.match_used) {
    const GkMatch& m = c.match[mid];
    if (matchgen && emit_match(o, c, mid)) continue;
    o << "#define M" << mid << " (GkMatch{" << hx(m.flags) << ", " << m.kinds_off << "u, " << m.kinds_n << "u, " << m.ns_off << "u, " << m.ns_n << "u, " << m.exns_off << "u, "
      << m.exns_n << "u, " << m.lsel_off << "u, " << m.lsel_n << "u, " << m.nssel_off << "u, " << m.nssel_n << "u, " << m.name_mode << "u, " << m.name_boff << "u, "
      << m.name_len << "u, 0u, 0u})\n";
    o << "GK_SPEC_FN int gk_spec_match_" << mid << "(const GkBatch& B, const uint32_t* __restrict__ pool, const uint8_t* __restrict__ cbytes, const uint32_t obj) {\n"
      << "  return gk_match(B, pool, cbytes, M" << mid << ", obj);\n}\n";
  }
  // returns true when the object has more rows in some scope than a mask holds: its words are then meaningless (the caller
  // hands the tile to the interpreter, which stores them again; a matcher error may be listed twice, with the same code)
  o << "\nGK_SPEC_FN bool gk_spec_object(const GkBatch& B, const GkColumn* cols, const GkScope* scopes, const uint32_t* __restrict__ pool, const uint8_t* __restrict__ cbytes,\n"
       "                               const uint32_t* actw, const GkOut& out, const uint32_t obj, uint32_t* vw, uint32_t* ew) {\n";
  o << "  const uint32_t lo0 = obj, hi0 = obj + 1u, n0 = 1u, f0 = 1u;\n  (void)lo0; (void)hi0; (void)n0; (void)f0; (void)out; (void)pool; (void)cbytes; (void)cols;\n";
  o << "  const bool skip = (GK_SPEC_LD(B.flags + obj) & GK_F_SKIP) != 0u;\n  (void)skip;\n  bool big = false;\n";
  for (size_t s = 1; s < NS; ++s) {
    if (!g.scope_used[s]) continue;
    const int P = c.schema.scopes[s].parent;
    o << "  const uint32_t* __restrict__ o" << s << " = scopes[" << s << "].off;\n";
    // (a scope with more rows than a mask holds counts as EMPTY from here on and raises `big`: no branch -- an early return here keeps
    // every other load of the object waiting behind the three dependent levels of CSR offsets: 0.71-0.73 vs 0.685 ms)
    o << "  const uint32_t lo" << s << " = GK_SPEC_LD(o" << s << " + lo" << P << "), r" << s << " = GK_SPEC_LD(o" << s << " + hi" << P << ") - lo" << s << ";\n";
    o << "  const uint32_t n" << s << " = r" << s << " > 32u ? 0u : r" << s << ", hi" << s << " = lo" << s << " + n" << s << ";\n";
    o << "  const uint32_t f" << s << " = n" << s << " >= 32u ? 0xffffffffu : ((1u << n" << s << ") - 1u);\n  (void)hi" << s << "; (void)f" << s << ";\n  big = big || r" << s
      << " > 32u;\n";
  }
  // ---- emission order of the constraints: by template kind, then by the scopes their cones touch
  std::vector<std::vector<uint32_t>> cone_atoms(g.nvar);   // variable -> atoms in its cone (memoised)
  std::vector<uint8_t> cone_done(g.nvar, 0);
  std::function<const std::vector<uint32_t>&(uint32_t)> cone = [&](uint32_t v) -> const std::vector<uint32_t>& {
    if (cone_done[v]) return cone_atoms[v];
    cone_done[v] = 1;
    std::set<uint32_t> acc;
    if (g.atom_of[v] >= 0) {
      acc.insert((uint32_t)g.atom_of[v]);
    } else if (g.producer[v] >= 0) {
      for (uint32_t d : g.nodes[(size_t)g.producer[v]].deps) {
        const std::vector<uint32_t>& sub = cone(d);
        acc.insert(sub.begin(), sub.end());
      }
    }
    cone_atoms[v].assign(acc.begin(), acc.end());
    return cone_atoms[v];
  };
  auto var_of = [&](uint32_t slot) -> int { return slot < g.cur.size() ? g.cur[slot] : -1; };
  struct Ord {
    std::string kind;
    std::vector<int> sig;
    uint32_t cix;
    bool operator<(const Ord& x) const { return std::tie(kind, sig, cix) < std::tie(x.kind, x.sig, x.cix); }
  };
  std::vector<Ord> order;
  for (uint32_t cix = 0; cix < C; ++cix) {
    const GkOutEnt& oe = c.outs[cix];
    Ord e;
    e.cix = cix;
    e.kind = cix < c.order.size() && c.order[cix] ? c.order[cix]->kind : std::string();
    if (!(oe.flags & 3u) && var_of(oe.prog_slot) >= 0) {
      std::set<int> sc;
      for (uint32_t ai : cone((uint32_t)var_of(oe.prog_slot))) sc.insert((int)g.atoms[ai].scope);
      e.sig.assign(sc.rbegin(), sc.rend());
    }
    order.push_back(std::move(e));
  }
  std::sort(order.begin(), order.end());
  // ---- atom groups: ONE row loop per scope computes every atom of that scope (~45 mask registers of the container scope alive at
  // once, some spilled).  A loop per (scope, template kind) keeps fewer masks alive and measured no better on B200 (0.726 vs 0.719 ms
  // at 1 M Pods x 50 constraints): every extra loop is one more exposed memory latency per object (profiles/experiments/README.md).
  std::map<uint32_t, int> group_ix;   // scope -> row loop
  std::vector<std::vector<uint32_t>> groups;
  for (const Ord& e : order) {
    const GkOutEnt& oe = c.outs[e.cix];
    if ((oe.flags & 3u) || var_of(oe.prog_slot) < 0) continue;
    for (uint32_t ai : cone((uint32_t)var_of(oe.prog_slot))) {
      Gen::AtomRec& r = g.atoms[ai];
      if (r.group >= 0) continue;
      auto it = group_ix.find(r.scope);
      if (it == group_ix.end()) {
        it = group_ix.emplace(r.scope, (int)groups.size()).first;
        groups.emplace_back();
      }
      r.group = it->second;
      groups[(size_t)r.group].push_back(ai);
    }
  }
  std::vector<uint8_t> group_done(groups.size(), 0);
  std::set<std::string> ptr_declared;
  auto emit_group = [&](int gi) {
    if (gi < 0 || group_done[(size_t)gi]) return;
    group_done[(size_t)gi] = 1;
    const std::vector<uint32_t>& members = groups[(size_t)gi];
    const uint32_t s = g.atoms[members[0]].scope;
    std::map<uint32_t, uint32_t> need;
    for (uint32_t ai : members)
      for (auto& kv : g.atoms[ai].need) need[kv.first] |= kv.second;
    o << "  // ---- atoms of scope " << s << " (row loop " << gi << ")\n";
    auto decl = [&](const std::string& name, const std::string& text) {
      if (ptr_declared.insert(name).second) o << text;
    };
    for (auto& kv : need) {
      const uint32_t ci = kv.first, enc = kv.second;
      const std::string K = std::to_string(ci);
      if (enc & GK_ENC_VT) decl("pt" + K, "  const uint8_t* __restrict__ pt" + K + " = cols[" + K + "].vt;\n");
      if (enc & GK_ENC_SID) decl("pi" + K, "  const uint32_t* __restrict__ pi" + K + " = cols[" + K + "].sid;\n");
      if (enc & GK_ENC_NUM) decl("pn" + K, "  const long long* __restrict__ pn" + K + " = reinterpret_cast<const long long*>(cols[" + K + "].num);\n");
      if (enc & GK_ENC_HEAD) decl("ph" + K, "  const uint4* __restrict__ ph" + K + " = reinterpret_cast<const uint4*>(cols[" + K + "].head);\n");
    }
    o << "  uint32_t";
    for (size_t k = 0; k < members.size(); ++k) o << (k ? ", v" : " v") << g.atoms[members[k]].var << " = 0u";
    o << ";\n";
    if (s == 0) o << "  {\n    const size_t row = obj;\n    const uint32_t bit = 1u;\n";
    else o << "  _Pragma(\"unroll 1\") for (uint32_t j = 0; j < n" << s << "; ++j) {\n    const size_t row = (size_t)lo" << s << " + j;\n    const uint32_t bit = 1u << j;\n    (void)bit;\n";
    for (auto& kv : need) {
      const uint32_t ci = kv.first, enc = kv.second;
      if (enc & GK_ENC_VT) o << "    const uint32_t t" << ci << " = GK_SPEC_LD(pt" << ci << " + row);\n";
      if (enc & GK_ENC_SID) o << "    const uint32_t i" << ci << " = GK_SPEC_LD(pi" << ci << " + row);\n";
      if (enc & GK_ENC_NUM) o << "    const long long n" << ci << " = GK_SPEC_LD(pn" << ci << " + row);\n";
      if (enc & GK_ENC_HEAD) o << "    const uint4 h" << ci << "a = GK_SPEC_LD(ph" << ci << " + 2 * row), h" << ci << "b = GK_SPEC_LD(ph" << ci << " + 2 * row + 1);\n";
    }
    for (uint32_t ai : members) {
      o << "    if (" << g.atoms[ai].expr << ") v" << g.atoms[ai].var << " |= bit;\n";
    }
    o << "  }\n";
  };
  // ---- everything else depth-first from the results: a node right after what it reads
  std::function<void(uint32_t)> emit_var = [&](uint32_t v) {
    if (g.atom_of[v] >= 0) {
      emit_group(g.atoms[(size_t)g.atom_of[v]].group);
      return;
    }
    const int pr = g.producer[v];
    if (pr < 0) return;
    SpecNode& n = g.nodes[(size_t)pr];
    if (n.done) return;
    n.done = true;
    for (uint32_t d : n.deps) emit_var(d);
    o << n.code;
  };
  for (const Ord& ent : order) {
    const uint32_t cix = ent.cix;
    const GkOutEnt& oe = c.outs[cix];
    g.deps.clear();
    const std::string pv = (oe.flags & 1u) ? "1u" : (oe.flags & 2u) ? "0u" : "(" + g.rd(oe.prog_slot) + " & 1u)";
    const std::string mt = g.rd(oe.match_slot), er = g.rd(oe.err_slot);
    for (uint32_t d : g.deps) emit_var(d);
    o << "  vw[" << (cix >> 5) << "] |= (" << pv << " & " << mt << ") << " << (cix & 31u) << ";   // constraint " << cix << " (" << ent.kind << ")\n  ew[" << (cix >> 5) << "] |= (" << er
      << " & 1u) << " << (cix & 31u) << ";\n";
  }
  // the enforcement-point filter: one packed word per 32 constraints (actw), not a test per constraint
  for (uint32_t w = 0; w < W; ++w) o << "  vw[" << w << "] &= actw[" << w << "];\n  ew[" << w << "] &= actw[" << w << "];\n";
  o << "  return big;\n}\n";
  o << kWrapper;
  out.src = o.str();
  out.words = W;
  auto r16 = [](size_t x) { return (x + 15) / 16 * 16; };
  out.smem = r16(c.schema.cols.size() * sizeof(GkColumn) + NS * sizeof(GkScope) + 3 * (size_t)W * 32 * 4 + (size_t)W * 4) + 16;
  out.n_fast = g.n_fast;
  out.n_generic = g.n_generic;
  return out;
}

}  // namespace gk
