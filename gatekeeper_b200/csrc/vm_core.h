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
    if (GK_LD(a + i) != GK_LD(b + i)) return false;
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
    if (GK_LD(s + i) == GK_LD(p) && gk_bytes_eq(s + i, p, pl)) return true;
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
    if (GK_LD(kv + 2 * i) == key) return GK_LD(kv + 2 * i + 1);
  return GK_NONE;
}

// labels.Selector.Matches over pool-encoded requirements [key, op, nvals, vals...]
GK_HD bool gk_selector(const uint32_t* pool, uint32_t off, uint32_t nreq, const uint32_t* kv, uint32_t lo, uint32_t hi) {
  for (uint32_t r = 0; r < nreq; ++r) {
    uint32_t key = GK_LD(pool + off), op = GK_LD(pool + off + 1), nv = GK_LD(pool + off + 2);
    uint32_t val = gk_label(kv, lo, hi, key);
    bool has = val != GK_NONE, in = false;
    if (has)
      for (uint32_t j = 0; j < nv; ++j) in = in || GK_LD(pool + off + 3 + j) == val;
    bool ok = op == GK_SEL_IN ? in : op == GK_SEL_NOTIN ? !in : op == GK_SEL_EXISTS ? has : !has;
    if (!ok) return false;
    off += 3 + nv;
  }
  return true;
}

// returns 1 match, 0 no match, <0 = -(GK_E_* code)
GK_HD int gk_match_row(const GkBatch& b, const uint32_t* pool, const uint8_t* cbytes, const GkMatch& m, uint32_t row,
                       uint32_t obj) {
  const uint32_t fl = GK_LD(b.flags + row);
  const bool is_ns = fl & GK_F_IS_NS;
  // 1 kinds -- match.go:181-201 (version ignored)
  if (m.kinds_n) {
    const uint32_t kind = GK_LD(b.kind_sid + row), group = GK_LD(b.group_sid + row);
    bool any = false;
    uint32_t off = m.kinds_off;
    for (uint32_t e = 0; e < m.kinds_n && !any; ++e) {
      uint32_t nk = GK_LD(pool + off), ng = GK_LD(pool + off + 1), wild = GK_LD(pool + off + 2);
      bool km = nk == 0 || (wild & 1), gm = ng == 0 || (wild & 2);
      for (uint32_t j = 0; j < nk && !km; ++j) km = GK_LD(pool + off + 3 + j) == kind;
      if (km)
        for (uint32_t j = 0; j < ng && !gm; ++j) gm = GK_LD(pool + off + 3 + nk + j) == group;
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
    if (fl & GK_F_NSNAME) {
      const uint32_t s0 = GK_LD(b.nsn_off + row);
      const uint8_t* s = b.nsn_bytes + s0;
      const uint32_t sl = GK_LD(b.nsn_off + row + 1) - s0;
      if (m.ns_n) {
        bool any = false;
        for (uint32_t j = 0; j < m.ns_n && !any; ++j) {
          const uint32_t* e = pool + m.ns_off + 3 * j;
          any = gk_wild(GK_LD(e), cbytes + GK_LD(e + 1), GK_LD(e + 2), s, sl);
        }
        if (!any) return 0;
      }
      for (uint32_t j = 0; j < m.exns_n; ++j) {
        const uint32_t* e = pool + m.exns_off + 3 * j;
        if (gk_wild(GK_LD(e), cbytes + GK_LD(e + 1), GK_LD(e + 2), s, sl)) return 0;
      }
    }
  }
  // 5 labelSelector -- match.go:103-116
  if (m.flags & GK_M_HAS_LSEL) {
    if (m.flags & GK_M_LSEL_INVALID) return -GK_E_LSEL_INVALID;
    if (!gk_selector(pool, m.lsel_off, m.lsel_n, b.lbl_kv, GK_LD(b.lbl_off + row), GK_LD(b.lbl_off + row + 1))) return 0;
  }
  // 6 namespaceSelector -- match.go:73-101
  if (m.flags & GK_M_HAS_NSSEL) {
    const bool ns_obj = fl & GK_F_NS_OBJ, obj_ns = fl & GK_F_HAS_NS;
    if (is_ns || ns_obj || obj_ns) {
      if (m.flags & GK_M_NSSEL_INVALID) return -GK_E_NSSEL_INVALID;
      if (is_ns) {
        if (!gk_selector(pool, m.nssel_off, m.nssel_n, b.lbl_kv, GK_LD(b.lbl_off + row), GK_LD(b.lbl_off + row + 1))) return 0;
      } else {
        if (!ns_obj) return -GK_E_NS_MISSING;
        const uint32_t nr = GK_LD(b.nsrow + obj);
        if (!gk_selector(pool, m.nssel_off, m.nssel_n, b.nsl_kv, GK_LD(b.nsl_off + nr), GK_LD(b.nsl_off + nr + 1))) return 0;
      }
    }
  }
  // 7 name -- match.go:203-212
  if (m.flags & GK_M_HAS_NAME) {
    const uint8_t* p = cbytes + m.name_boff;
    const uint32_t a = GK_LD(b.name_off + row), a1 = GK_LD(b.name_off + row + 1);
    bool ok = gk_wild(m.name_mode, p, m.name_len, b.name_bytes + a, a1 - a);
    if (!ok) {
      const uint32_t g = GK_LD(b.gen_off + row), g1 = GK_LD(b.gen_off + row + 1);
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
  if (GK_LD(b.flags + obj) & GK_F_HAS_OBJ) {
    int r = gk_match_row(b, pool, cbytes, m, obj, obj);
    if (r) return r;
  } else {
    ++nil;
  }
  if (b.has_old && (GK_LD(b.flags + b.n + obj) & GK_F_HAS_OBJ)) {
    int r = gk_match_row(b, pool, cbytes, m, b.n + obj, obj);
    if (r < 0) return r - GK_E_FROM_OLD;   // the error text names the object that failed (matcher.go:58-60): here the old one
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

// One atom on one row.  Shared by the CUDA tile executor and the test-only host emulation.
GK_HD bool gk_atom(const GkColumn& c, uint32_t row, uint32_t op, uint32_t w2, uint32_t w3, const uint32_t* pool, const uint8_t* cbytes) {
  switch (op) {
    case GK_OP_TRUTHY: { const uint32_t vt = c.vt[row]; return vt != GK_VT_UNDEF && vt != GK_VT_FALSE; }
    case GK_OP_DEFINED: return c.vt[row] != GK_VT_UNDEF;
    case GK_OP_VTMASK: return ((1u << c.vt[row]) & w2) != 0;
    case GK_OP_SID_EQ: return c.sid[row] == w2;
    case GK_OP_SID_IN: {
      const uint32_t v = c.sid[row];
      uint32_t lo = w2, hi = w2 + w3;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1, x = pool[mid];
        if (x == v) return true;
        if (x < v) lo = mid + 1; else hi = mid;
      }
      return false;
    }
    case GK_OP_NUM_CMP: {
      const uint32_t vt = c.vt[row];
      if (vt == GK_VT_UNDEF || vt == GK_VT_NUM_INEXACT) return false;   // inexact numbers never reach the device: the flattener rejects the object
      if (vt == GK_VT_NUM) {
        const int64_t v = c.num[row];
        const int64_t k = (int64_t)(((uint64_t)pool[w2 + 1] << 32) | pool[w2]);
        return gk_cmp_apply(w3, v < k ? -1 : (v > k ? 1 : 0));
      }
      return gk_cmp_apply(w3, gk_vt_rank(vt) < 2 ? -1 : 1);
    }
    case GK_OP_SUFFIX:
    case GK_OP_CONTAINS: {
      if (c.vt[row] != GK_VT_STR) return false;
      const uint32_t a = c.boff[row], sl = c.boff[row + 1] - a;
      const uint8_t* s = c.bytes + a;
      return op == GK_OP_SUFFIX ? gk_suffix(s, sl, cbytes + w2, w3) : gk_contains(s, sl, cbytes + w2, w3);
    }
    case GK_OP_ANYPREFIX: {
      // prefix tests run on the fixed-width HEAD record: one aligned 32-byte load per row, word compares under a
      // length mask; only prefixes longer than 31 bytes continue in the byte pool
      if (c.vt[row] != GK_VT_STR) return false;
      uint32_t h[GK_HEAD_WORDS];
      const uint32_t* hp = c.head + (size_t)row * GK_HEAD_WORDS;
      for (int i = 0; i < GK_HEAD_WORDS; ++i) h[i] = hp[i];
      const uint32_t lenb = h[GK_HEAD_WORDS - 1] >> 24;   // min(len, 255)
      for (uint32_t j = 0; j < w3; ++j) {
        const uint32_t* e = pool + w2 + (size_t)j * GK_PREFIX_ENT;
        const uint32_t L = e[0];
        const uint32_t Lh = L < GK_HEAD_BYTES ? L : GK_HEAD_BYTES;
        if (lenb < Lh) continue;
        bool ok = true;
        for (uint32_t w = 0; w < GK_HEAD_WORDS; ++w) ok = ok && (((h[w] ^ e[2 + w]) & e[2 + GK_HEAD_WORDS + w]) == 0u);
        if (ok && L > GK_HEAD_BYTES) {
          const uint32_t a = c.boff[row], sl = c.boff[row + 1] - a;
          ok = gk_prefix(c.bytes + a, sl, cbytes + e[1], L);
        }
        if (ok) return true;
      }
      return false;
    }
    case GK_OP_ANYSUFFIX: {
      if (c.vt[row] != GK_VT_STR) return false;
      const uint32_t a = c.boff[row], sl = c.boff[row + 1] - a;
      const uint8_t* s = c.bytes + a;
      for (uint32_t j = 0; j < w3; ++j) {
        const uint32_t po = pool[w2 + 2 * j], pl = pool[w2 + 2 * j + 1];
        if (gk_suffix(s, sl, cbytes + po, pl)) return true;
      }
      return false;
    }
    default: return false;
  }
}
