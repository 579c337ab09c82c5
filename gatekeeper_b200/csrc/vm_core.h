// Per-(object, constraint) evaluation core: the spec.match pre-filter and the predicate interpreter.
// Written once as host/device inline code: kernels.cu runs it one thread per object on the GPU (the
// product path); tests/_hostemu compiles the very same functions for CPU-only unit tests of the lowering
// in the authoring container, which has no GPU.  Nothing in the product library calls it on the host.
//
// Semantics restated from the reference (file:line relative to /root/reference):
//   gk_match_row   : match.Matches            pkg/mutation/match/match.go:32-65 (8 matchers, fixed order, early exit)
//   gk_wild        : wildcard.Wildcard.Matches pkg/wildcard/wildcard.go:17-29
//   gk_wild_gen    : MatchesGenerateName       pkg/wildcard/wildcard.go:31-41
//   gk_match       : Matcher.Match / matchAny  pkg/target/matcher.go:21-71 (object OR oldObject)
#pragma once
#include "program.h"

GK_HD bool gk_bytes_eq(const uint8_t* a, const uint8_t* b, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i)
    if (a[i] != b[i]) return false;
  return true;
}

GK_HD bool gk_prefix(const uint8_t* s, uint32_t sl, const uint8_t* p, uint32_t pl) {
  return sl >= pl && gk_bytes_eq(s, p, pl);
}
GK_HD bool gk_suffix(const uint8_t* s, uint32_t sl, const uint8_t* p, uint32_t pl) {
  return sl >= pl && gk_bytes_eq(s + (sl - pl), p, pl);
}
GK_HD bool gk_contains(const uint8_t* s, uint32_t sl, const uint8_t* p, uint32_t pl) {
  if (pl == 0) return true;
  if (sl < pl) return false;
  for (uint32_t i = 0; i + pl <= sl; ++i)
    if (s[i] == p[0] && gk_bytes_eq(s + i, p, pl)) return true;
  return false;
}

// mode already strips the '*'s: EXACT "x", PREFIX "x*", SUFFIX "*x", CONTAINS "*x*"
GK_HD bool gk_wild(uint32_t mode, const uint8_t* p, uint32_t pl, const uint8_t* s, uint32_t sl) {
  switch (mode) {
    case GK_W_PREFIX: return gk_prefix(s, sl, p, pl);
    case GK_W_SUFFIX: return gk_suffix(s, sl, p, pl);
    case GK_W_CONTAINS: return gk_contains(s, sl, p, pl);
    default: return sl == pl && gk_bytes_eq(s, p, pl);
  }
}
GK_HD bool gk_wild_gen(uint32_t mode, const uint8_t* p, uint32_t pl, const uint8_t* s, uint32_t sl) {
  switch (mode) {
    case GK_W_PREFIX: return gk_prefix(s, sl, p, pl);
    case GK_W_CONTAINS: return gk_contains(s, sl, p, pl);
    default: return false;   // exact and "*x" never match a generateName
  }
}

// label lookup in a (key sid, value sid) run; returns value sid or GK_NONE
GK_HD uint32_t gk_label(const uint32_t* kv, uint32_t lo, uint32_t hi, uint32_t key) {
  for (uint32_t i = lo; i < hi; ++i)
    if (kv[2 * i] == key) return kv[2 * i + 1];
  return GK_NONE;
}

// labels.Selector.Matches over pool-encoded requirements [key, op, nvals, vals...]
GK_HD bool gk_selector(const uint32_t* pool, uint32_t off, uint32_t nreq, const uint32_t* kv, uint32_t lo, uint32_t hi) {
  for (uint32_t r = 0; r < nreq; ++r) {
    uint32_t key = pool[off], op = pool[off + 1], nv = pool[off + 2];
    uint32_t val = gk_label(kv, lo, hi, key);
    bool has = val != GK_NONE, in = false;
    if (has)
      for (uint32_t j = 0; j < nv; ++j) in = in || pool[off + 3 + j] == val;
    bool ok = op == GK_SEL_IN ? in : op == GK_SEL_NOTIN ? !in : op == GK_SEL_EXISTS ? has : !has;
    if (!ok) return false;
    off += 3 + nv;
  }
  return true;
}

// returns 1 match, 0 no match, <0 = -(GK_E_* code)
GK_HD int gk_match_row(const GkBatch& b, const uint32_t* pool, const uint8_t* cbytes, const GkMatch& m, uint32_t row,
                       uint32_t obj) {
  const uint32_t fl = b.flags[row];
  const bool is_ns = fl & GK_F_IS_NS;
  // 1 kinds -- match.go:181-201 (version ignored)
  if (m.kinds_n) {
    const uint32_t kind = b.kind_sid[row], group = b.group_sid[row];
    bool any = false;
    uint32_t off = m.kinds_off;
    for (uint32_t e = 0; e < m.kinds_n && !any; ++e) {
      uint32_t nk = pool[off], ng = pool[off + 1], wild = pool[off + 2];
      bool km = nk == 0 || (wild & 1), gm = ng == 0 || (wild & 2);
      for (uint32_t j = 0; j < nk && !km; ++j) km = pool[off + 3 + j] == kind;
      if (km)
        for (uint32_t j = 0; j < ng && !gm; ++j) gm = pool[off + 3 + nk + j] == group;
      any = km && gm;
      off += 3 + nk + ng;
    }
    if (!any) return 0;
  }
  // 2 scope -- match.go:214-227
  {
    const bool has_ns = fl & (GK_F_HAS_NS | GK_F_NS_OBJ);
    if ((m.flags & GK_M_SCOPE_CLUSTER) && !(is_ns || !has_ns)) return 0;
    if ((m.flags & GK_M_SCOPE_NAMESPACED) && !(!is_ns && has_ns)) return 0;
  }
  // 3/4 namespaces, excludedNamespaces -- match.go:118-179
  if (m.ns_n || m.exns_n) {
    const uint32_t sid = b.nsname_sid[row];
    if (sid != GK_NONE) {
      const uint8_t* s = b.dict_bytes + b.dict_off[sid] + 1;   // skip the intern type char
      const uint32_t sl = b.dict_off[sid + 1] - b.dict_off[sid] - 1;
      if (m.ns_n) {
        bool any = false;
        for (uint32_t j = 0; j < m.ns_n && !any; ++j) {
          const uint32_t* e = pool + m.ns_off + 3 * j;
          any = gk_wild(e[0], cbytes + e[1], e[2], s, sl);
        }
        if (!any) return 0;
      }
      for (uint32_t j = 0; j < m.exns_n; ++j) {
        const uint32_t* e = pool + m.exns_off + 3 * j;
        if (gk_wild(e[0], cbytes + e[1], e[2], s, sl)) return 0;
      }
    }
  }
  // 5 labelSelector -- match.go:103-116
  if (m.flags & GK_M_HAS_LSEL) {
    if (m.flags & GK_M_LSEL_INVALID) return -GK_E_LSEL_INVALID;
    if (!gk_selector(pool, m.lsel_off, m.lsel_n, b.lbl_kv, b.lbl_off[row], b.lbl_off[row + 1])) return 0;
  }
  // 6 namespaceSelector -- match.go:73-101
  if (m.flags & GK_M_HAS_NSSEL) {
    const bool ns_obj = fl & GK_F_NS_OBJ, obj_ns = fl & GK_F_HAS_NS;
    if (is_ns || ns_obj || obj_ns) {
      if (m.flags & GK_M_NSSEL_INVALID) return -GK_E_NSSEL_INVALID;
      if (is_ns) {
        if (!gk_selector(pool, m.nssel_off, m.nssel_n, b.lbl_kv, b.lbl_off[row], b.lbl_off[row + 1])) return 0;
      } else {
        if (!ns_obj) return -GK_E_NS_MISSING;
        const uint32_t nr = b.nsrow[obj];
        if (!gk_selector(pool, m.nssel_off, m.nssel_n, b.nsl_kv, b.nsl_off[nr], b.nsl_off[nr + 1])) return 0;
      }
    }
  }
  // 7 name -- match.go:203-212
  if (m.flags & GK_M_HAS_NAME) {
    const uint8_t* p = cbytes + m.name_boff;
    const uint32_t a = b.name_off[row], a1 = b.name_off[row + 1];
    bool ok = gk_wild(m.name_mode, p, m.name_len, b.name_bytes + a, a1 - a);
    if (!ok) {
      const uint32_t g = b.gen_off[row], g1 = b.gen_off[row + 1];
      ok = gk_wild_gen(m.name_mode, p, m.name_len, b.gen_bytes + g, g1 - g);
    }
    if (!ok) return 0;
  }
  // 8 source -- match.go:229-253
  {
    if (m.flags & GK_M_SRC_INVALID) return -GK_E_SRC_INVALID_MATCH;
    const uint32_t msrc = (m.flags >> GK_M_SRC_SHIFT) & 7u, tsrc = (fl & GK_F_SRC_MASK) >> GK_F_SRC_SHIFT;
    if (tsrc == GK_SRC_EMPTY && msrc != GK_SRC_ALL) return -GK_E_SRC_UNSPECIFIED;
    if (msrc != GK_SRC_ALL) {
      if (tsrc == GK_SRC_INVALID) return -GK_E_SRC_INVALID_OBJ;
      if (msrc != tsrc) return 0;
    }
  }
  return 1;
}

GK_HD int gk_match(const GkBatch& b, const uint32_t* pool, const uint8_t* cbytes, const GkMatch& m, uint32_t obj) {
  if (!(m.flags & GK_M_HAS_MATCH)) return 1;   // matcher.go:22-25
  int nil = 0;
  if (b.flags[obj] & GK_F_HAS_OBJ) {
    int r = gk_match_row(b, pool, cbytes, m, obj, obj);
    if (r) return r;
  } else {
    ++nil;
  }
  if (b.has_old && (b.flags[b.n + obj] & GK_F_HAS_OBJ)) {
    int r = gk_match_row(b, pool, cbytes, m, b.n + obj, obj);
    if (r) return r;
  } else {
    ++nil;
  }
  return nil == 2 ? -GK_E_NO_OBJECT : 0;
}

GK_HD int gk_vt_rank(uint32_t vt) {
  // null < bool < number < string < array < object < set
  return vt == GK_VT_NULL ? 0 : (vt == GK_VT_FALSE || vt == GK_VT_TRUE) ? 1 : (vt == GK_VT_NUM || vt == GK_VT_NUM_INEXACT) ? 2
         : vt == GK_VT_STR ? 3 : vt == GK_VT_ARR ? 4 : vt == GK_VT_OBJ ? 5 : 6;
}

GK_HD bool gk_cmp_apply(uint32_t op, int c) {
  switch (op) {
    case GK_CMP_LT: return c < 0;
    case GK_CMP_LE: return c <= 0;
    case GK_CMP_GT: return c > 0;
    case GK_CMP_GE: return c >= 0;
    case GK_CMP_EQ: return c == 0;
    default: return c != 0;
  }
}

// ---- warp-uniform helpers: on the GPU every lane of a warp executes the same instruction stream; the host
// emulation is a "warp" of one lane.
#ifdef __CUDA_ARCH__
#define GK_WARP_MAX(x) __reduce_max_sync(0xffffffffu, (x))
#else
#define GK_WARP_MAX(x) (x)
#endif

struct GkLoop {
  uint32_t it, end;
};

// Evaluates the postfix predicate starting at `pc` for object `obj`.  MUST be called by all lanes of the warp
// with the same pc (lanes without a live object pass live = false and evaluate to false).  `cse` carries the
// shared sub-formula bits of this object across the constraints of the batch; *flag is set to GK_E_NUM_RANGE
// when an ordered compare met a number the flattener could not represent exactly as int64.
GK_HD bool gk_eval_prog(const GkColumn* cols, const GkScope* scopes, const GkInstr* instr, const uint32_t* pool, const uint8_t* cbytes,
                        uint32_t pc, uint32_t obj, bool live, unsigned long long& cse, unsigned long long& cse_valid, int* flag) {
  unsigned long long st = 0;
  uint32_t it1 = 0, it2 = 0, it3 = 0, it4 = 0, en1 = 0, en2 = 0, en3 = 0, en4 = 0;
  uint32_t tr1 = 0, tr2 = 0, tr3 = 0, tr4 = 0;   // warp-uniform remaining trip counts
  for (;;) {
    const GkInstr in = instr[pc];
    const uint32_t op = in.w0 & 0xffu, slot = (in.w0 >> 8) & 0xffu, col = in.w0 >> 16;
    switch (op) {
      case GK_OP_END: return (st & 1ull) != 0;
      case GK_OP_AND: {
        const unsigned long long t = st & 1ull;
        st >>= 1;
        st &= (t | ~1ull);
        break;
      }
      case GK_OP_OR: {
        const unsigned long long t = st & 1ull;
        st >>= 1;
        st |= t;
        break;
      }
      case GK_OP_NOT: st ^= 1ull; break;
      case GK_OP_PUSH: st = (st << 1) | (in.w1 & 1u); break;
      case GK_OP_CSE_TRY:
        if ((cse_valid >> in.w1) & 1ull) {   // warp-uniform
          st = (st << 1) | ((cse >> in.w1) & 1ull);
          pc = in.w2;
          continue;
        }
        break;
      case GK_OP_CSE_STORE:
        cse = (cse & ~(1ull << in.w1)) | ((st & 1ull) << in.w1);
        cse_valid |= 1ull << in.w1;
        break;
      case GK_OP_LOOP_BEGIN: {
        // parent row (col field = parent slot) and whether this lane is still inside the parent loop
        const uint32_t prow = col == 0 ? obj : col == 1 ? it1 : col == 2 ? it2 : it3;
        const bool pvalid = col == 0 ? live : col == 1 ? it1 < en1 : col == 2 ? it2 < en2 : it3 < en3;
        uint32_t lo = 0, hi = 0;
        if (pvalid) {
          const uint32_t* off = scopes[in.w1].off;
          lo = off[prow];
          hi = off[prow + 1];
        }
        const uint32_t trip = GK_WARP_MAX(hi - lo);
        if (slot == 1) { it1 = lo; en1 = hi; tr1 = trip; }
        else if (slot == 2) { it2 = lo; en2 = hi; tr2 = trip; }
        else if (slot == 3) { it3 = lo; en3 = hi; tr3 = trip; }
        else { it4 = lo; en4 = hi; tr4 = trip; }
        st <<= 1;   // accumulator = false
        if (trip == 0) {
          pc = in.w2;
          continue;
        }
        break;
      }
      case GK_OP_LOOP_END: {
        bool valid;
        uint32_t left;
        if (slot == 1) { valid = it1 < en1; ++it1; left = --tr1; }
        else if (slot == 2) { valid = it2 < en2; ++it2; left = --tr2; }
        else if (slot == 3) { valid = it3 < en3; ++it3; left = --tr3; }
        else { valid = it4 < en4; ++it4; left = --tr4; }
        const unsigned long long t = (st & 1ull) & (valid ? 1ull : 0ull);
        st >>= 1;
        st |= t;
        if (left) {
          pc = in.w2;
          continue;
        }
        break;
      }
      default: {
        // ---- atoms: push one bit
        const uint32_t row = slot == 0 ? obj : slot == 1 ? it1 : slot == 2 ? it2 : slot == 3 ? it3 : it4;
        const bool valid = slot == 0 ? live : slot == 1 ? it1 < en1 : slot == 2 ? it2 < en2 : slot == 3 ? it3 < en3 : it4 < en4;
        bool r = false;
        if (valid) {
          const GkColumn& c = cols[col];
          switch (op) {
            case GK_OP_TRUTHY: { const uint32_t vt = c.vt[row]; r = vt != GK_VT_UNDEF && vt != GK_VT_FALSE; break; }
            case GK_OP_DEFINED: r = c.vt[row] != GK_VT_UNDEF; break;
            case GK_OP_VTMASK: r = ((1u << c.vt[row]) & in.w1) != 0; break;
            case GK_OP_SID_EQ: r = c.sid[row] == in.w1; break;
            case GK_OP_SID_IN: {
              const uint32_t v = c.sid[row];
              uint32_t lo = in.w1, hi = in.w1 + in.w3;
              while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1, x = pool[mid];
                if (x == v) { r = true; break; }
                if (x < v) lo = mid + 1; else hi = mid;
              }
              break;
            }
            case GK_OP_NUM_CMP: {
              const uint32_t vt = c.vt[row];
              if (vt == GK_VT_UNDEF) break;
              if (vt == GK_VT_NUM) {
                const int64_t v = c.num[row];
                const int64_t k = (int64_t)(((uint64_t)pool[in.w1 + 1] << 32) | pool[in.w1]);
                r = gk_cmp_apply(in.w3, v < k ? -1 : (v > k ? 1 : 0));
              } else if (vt == GK_VT_NUM_INEXACT) {
                *flag = GK_E_NUM_RANGE;
              } else {
                r = gk_cmp_apply(in.w3, gk_vt_rank(vt) < 2 ? -1 : 1);
              }
              break;
            }
            case GK_OP_PREFIX:
            case GK_OP_SUFFIX:
            case GK_OP_CONTAINS: {
              if (c.vt[row] != GK_VT_STR) break;
              const uint32_t a = c.boff[row], sl = c.boff[row + 1] - a;
              const uint8_t* s = c.bytes + a;
              r = op == GK_OP_PREFIX   ? gk_prefix(s, sl, cbytes + in.w1, in.w3)
                  : op == GK_OP_SUFFIX ? gk_suffix(s, sl, cbytes + in.w1, in.w3)
                                       : gk_contains(s, sl, cbytes + in.w1, in.w3);
              break;
            }
            case GK_OP_ANYPREFIX:
            case GK_OP_ANYSUFFIX: {
              if (c.vt[row] != GK_VT_STR) break;
              const uint32_t a = c.boff[row], sl = c.boff[row + 1] - a;
              const uint8_t* s = c.bytes + a;
              for (uint32_t j = 0; j < in.w3 && !r; ++j) {
                const uint32_t po = pool[in.w1 + 2 * j], pl = pool[in.w1 + 2 * j + 1];
                r = op == GK_OP_ANYPREFIX ? gk_prefix(s, sl, cbytes + po, pl) : gk_suffix(s, sl, cbytes + po, pl);
              }
              break;
            }
            default: break;
          }
        }
        st = (st << 1) | (r ? 1ull : 0ull);
      }
    }
    ++pc;
  }
}
