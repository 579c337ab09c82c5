// Extraction program builder: compiled Schema (closures classified by lower.cpp's closure_xinfo) -> the flat tables the
// ingest kernels interpret (ingest_core.h GkXProg).
#include "xprog.hpp"

#include <cstdlib>
#include <map>

#include "engine.hpp"

namespace gk {

namespace {

struct Builder {
  XProgHost& x;
  std::map<std::string, uint32_t> ix;   // closure key -> entry

  uint32_t add_bytes(const std::string& s) {
    uint32_t off = (uint32_t)x.xbytes.size();
    x.xbytes.insert(x.xbytes.end(), s.begin(), s.end());
    return off;
  }
  // keys -> (is_index, byte_off | index, len) triples; false when a key can never match a JSON member / element
  bool put_keys(const std::vector<VP>& keys, size_t from, uint32_t& off, uint32_t& n) {
    off = (uint32_t)x.xkeys.size();
    n = 0;
    for (size_t i = from; i < keys.size(); ++i) {
      const VP& k = keys[i];
      if (k->t == VT::Str) {
        x.xkeys.push_back(0);
        x.xkeys.push_back(add_bytes(k->s));
        x.xkeys.push_back((uint32_t)k->s.size());
      } else if (k->t == VT::Num) {
        int64_t v;
        if (!num_fits_i64(k->n, &v) || v < 0 || v > 0x7fffffff) return false;
        x.xkeys.push_back(1);
        x.xkeys.push_back((uint32_t)v);
        x.xkeys.push_back(0);
      } else {
        return false;
      }
      ++n;
    }
    return true;
  }
  static uint32_t root_of(const std::string& k) {
    static const std::pair<const char*, uint32_t> kRoots[] = {
        {"object", GK_R_OBJECT},       {"kind", GK_R_KIND},           {"name", GK_R_NAME},       {"namespace", GK_R_NAMESPACE},
        {"oldObject", GK_R_OLDOBJECT}, {"operation", GK_R_OPERATION}, {"uid", GK_R_UID},         {"options", GK_R_OPTIONS},
        {"resource", GK_R_RESOURCE},   {"userInfo", GK_R_USERINFO}};
    for (auto& r : kRoots)
      if (k == r.first) return r.second;
    return GK_R_UNDEF;   // namespaceObject (not part of a blob review), dryRun, requestKind ...: absent from the envelope
  }
  // a path rooted at `input` (keys[0] == "review")
  uint32_t add_input_path(const std::vector<VP>& keys) {
    GkXClosure c{};
    c.kind = GK_X_PATH;
    c.base = -1;
    if (keys.size() == 1) {
      c.root = GK_R_REVIEW;
    } else if (keys[1]->t != VT::Str) {
      c.root = GK_R_UNDEF;
    } else {
      c.root = root_of(keys[1]->s);
      if (!put_keys(keys, 2, c.keys_off, c.nkeys)) c.root = GK_R_UNDEF, c.nkeys = 0;
    }
    x.cl.push_back(c);
    x.cl_src.push_back(nullptr);
    x.cl_args.emplace_back();
    return (uint32_t)x.cl.size() - 1;
  }
  // A path from a base closure is a CHAIN of one-key steps, each step shared by every path that goes through it
  // (`resources`, `resources.limits`, `resources.limits.cpu`, ... look `resources` up once per row: ingest_core.h memoises
  // a row's steps).  Beyond 10 keys the rest stays one closure (gk_x_eval walks chains of at most 16).
  std::map<std::string, uint32_t> step_ix;   // "<base closure>/<key>" -> step closure
  uint32_t add_step(uint32_t base, const std::vector<VP>& keys) {
    GkXClosure c{};
    c.kind = GK_X_PATH;
    c.base = (int32_t)base;
    if (!put_keys(keys, 0, c.keys_off, c.nkeys)) {   // a key that matches nothing: an always-undefined path
      c.base = -1;
      c.root = GK_R_UNDEF;
      c.nkeys = 0;
    }
    x.cl.push_back(c);
    x.cl_src.push_back(nullptr);
    x.cl_args.emplace_back();
    return (uint32_t)x.cl.size() - 1;
  }
  uint32_t add_path(uint32_t base, const std::vector<VP>& keys) {
    uint32_t cur = base;
    for (size_t i = 0; i < keys.size(); ++i) {
      if (i >= 10) return add_step(cur, std::vector<VP>(keys.begin() + (long)i, keys.end()));
      const std::string k = std::to_string(cur) + "/" + intern_key(keys[i]);
      auto it = step_ix.find(k);
      if (it == step_ix.end()) it = step_ix.emplace(k, add_step(cur, {keys[i]})).first;
      cur = it->second;
    }
    return cur;
  }

  // ---- native programs for closures built from slicing string builtins over one leaf (ingest_core.h GK_SX_*)
  struct SxOp {
    uint32_t op, a, b;
  };
  static const CapArg* cap_of(const Closure& c, int vid) {
    for (auto& cp : c.caps)
      if (cp.first == vid) return &cp.second;
    return nullptr;
  }
  static bool ascii(const std::string& s2) {
    for (unsigned char ch : s2)
      if (ch >= 0x80) return false;
    return true;
  }
  static int arg_index(const std::vector<XInfo::Arg>& args, const XInfo::Arg& a) {
    for (size_t i = 0; i < args.size(); ++i) {
      if ((bool)args[i].leaf != (bool)a.leaf || (a.leaf && args[i].leaf->key != a.leaf->key) || args[i].keys.size() != a.keys.size()) continue;
      bool same = true;
      for (size_t k = 0; k < a.keys.size() && same; ++k) same = v_eq(args[i].keys[k], a.keys[k]);
      if (same) return (int)i;
    }
    return -1;
  }
  // value of closure `cl` (then indexed by literal `keys`) as ops
  bool sx_closure(const CP& cl, const std::vector<VP>& keys, const std::vector<XInfo::Arg>& args, std::vector<SxOp>& prog, int depth) {
    if (depth > 6) return false;
    const XInfo xi = closure_xinfo(*cl);
    if (xi.k == XK::Path || xi.k == XK::Elem || xi.k == XK::Key) {
      std::vector<XInfo::Arg> one;
      if (!xinfo_leaf_args(cl, keys, one) || one.size() != 1) return false;
      const int k = arg_index(args, one[0]);
      if (k < 0) return false;
      prog.push_back({GK_SX_LEAF, (uint32_t)k, 0});
      return true;
    }
    if (xi.k != XK::Lut || !cl->term) return false;
    if (!sx_term(*cl, cl->term.get(), args, prog, depth + 1)) return false;
    for (auto& k : keys) {
      int64_t v;
      if (k->t != VT::Num || !num_fits_i64(k->n, &v) || v < 0 || v > 0xffff) return false;
      prog.push_back({GK_SX_INDEX, (uint32_t)v, 0});
    }
    return true;
  }
  static bool is_last_index(const Term& a, int head_vid) {   // minus(count(<head>), 1)
    if (a.k != TK::Call || a.name != "minus" || a.args.size() != 2) return false;
    const Term& c = *a.args[0];
    const Term& one = *a.args[1];
    int64_t v;
    if (one.k != TK::Scalar || one.val->t != VT::Num || !num_fits_i64(one.val->n, &v) || v != 1) return false;
    return c.k == TK::Call && c.name == "count" && c.args.size() == 1 && c.args[0]->k == TK::Var && c.args[0]->vid == head_vid;
  }
  bool sx_term(const Closure& c, const Term* t, const std::vector<XInfo::Arg>& args, std::vector<SxOp>& prog, int depth) {
    if (!t || depth > 6) return false;
    const Module& m = *c.mod;
    switch (t->k) {
      case TK::Var: {
        const CapArg* cap = cap_of(c, t->vid);
        if (!cap || cap->k != CapArg::Col) return false;
        return sx_closure(cap->col, {}, args, prog, depth + 1);
      }
      case TK::Ref: {
        if (t->head->k != TK::Var) return false;
        const CapArg* cap = cap_of(c, t->head->vid);
        if (!cap || cap->k != CapArg::Col) return false;
        std::vector<VP> keys;
        size_t i = 0;
        for (; i < t->args.size() && t->args[i]->k == TK::Scalar; ++i) keys.push_back(t->args[i]->val);
        if (!sx_closure(cap->col, keys, args, prog, depth + 1)) return false;
        for (; i < t->args.size(); ++i) {
          if (is_last_index(*t->args[i], t->head->vid) && keys.empty() && i == 0) prog.push_back({GK_SX_LAST, 0, 0});
          else return false;
        }
        return true;
      }
      case TK::Call: {
        if (m.is_rule(t->name)) return false;
        if ((t->name == "split" || t->name == "trim") && t->args.size() == 2 && t->args[1]->k == TK::Scalar && t->args[1]->val->t == VT::Str) {
          const std::string& lit = t->args[1]->val->s;
          if (!ascii(lit) || (t->name == "split" && lit.empty())) return false;
          if (!sx_term(c, t->args[0].get(), args, prog, depth + 1)) return false;
          prog.push_back({t->name == "split" ? (uint32_t)GK_SX_SPLIT : (uint32_t)GK_SX_TRIM, add_bytes(lit), (uint32_t)lit.size()});
          return true;
        }
        if (t->name == "count" && t->args.size() == 1) {
          if (!sx_term(c, t->args[0].get(), args, prog, depth + 1)) return false;
          prog.push_back({GK_SX_COUNT, 0, 0});
          return true;
        }
        return false;
      }
      default: return false;
    }
  }

  uint32_t add(const CP& cp) {
    auto it = ix.find(cp->key);
    if (it != ix.end()) return it->second;
    const XInfo xi = closure_xinfo(*cp);
    uint32_t id = 0;
    switch (xi.k) {
      case XK::Elem:
      case XK::Key: {
        GkXClosure c{};
        c.kind = xi.k == XK::Elem ? GK_X_ELEM : GK_X_KEY;
        c.scope = (uint32_t)cp->scope;
        x.cl.push_back(c);
        x.cl_src.push_back(cp);
        x.cl_args.emplace_back();
        id = (uint32_t)x.cl.size() - 1;
        break;
      }
      case XK::Path: {
        id = xi.from_input ? add_input_path(xi.keys) : add_path(add(xi.base), xi.keys);
        // (add_path returns the base itself for an empty key list: give the closure its own entry then)
        if (!xi.from_input && xi.keys.empty()) {
          GkXClosure c = x.cl[id];
          x.cl.push_back(c);
          x.cl_src.push_back(cp);
          x.cl_args.emplace_back();
          id = (uint32_t)x.cl.size() - 1;
        }
        x.cl_src[id] = cp;
        break;
      }
      case XK::Count: {
        const uint32_t b = add(xi.base);
        GkXClosure c{};
        c.kind = GK_X_COUNT;
        c.base = (int32_t)b;
        x.cl.push_back(c);
        x.cl_src.push_back(cp);
        x.cl_args.push_back(xi.args);
        id = (uint32_t)x.cl.size() - 1;
        break;
      }
      case XK::Lut: {
        std::vector<uint32_t> args;
        for (auto& a : xi.args) args.push_back(a.leaf ? add_path(add(a.leaf), a.keys) : add_input_path(a.keys));
        if (args.size() > GK_LUT_MAX_ARGS) throw RegoError{"rego_unsupported: device-ingest: lookup closure with too many arguments: " + cp->key};
        GkXClosure c{};
        c.kind = GK_X_LUT;
        c.base = -1;
        c.scope = (uint32_t)cp->scope;
        c.args_off = (uint32_t)x.xargs.size();
        c.nargs = (uint32_t)args.size();
        x.xargs.insert(x.xargs.end(), args.begin(), args.end());
        c.seed = xhash(GK_HASH_INIT, cp->key.data(), cp->key.size());
        {
          std::vector<SxOp> prog;
          static const bool no_sx = getenv("GK_NO_NATIVE_STRINGS") != nullptr;
          if (!no_sx && cp->term && sx_term(*cp, cp->term.get(), xi.args, prog, 0) && !prog.empty() && prog.size() <= 16) {
            c.sx_off = (uint32_t)x.xkeys.size();
            c.sx_n = (uint32_t)prog.size();
            for (auto& o : prog) {
              x.xkeys.push_back(o.op);
              x.xkeys.push_back(o.a);
              x.xkeys.push_back(o.b);
            }
          }
        }
        x.cl.push_back(c);
        x.cl_src.push_back(cp);
        x.cl_args.push_back(xi.args);
        id = (uint32_t)x.cl.size() - 1;
        break;
      }
      default: throw RegoError{"rego_unsupported: device-ingest: the ingest kernels cannot compute " + cp->key};
    }
    ix[cp->key] = id;
    return id;
  }
};

}  // namespace

std::shared_ptr<const XProgHost> build_xprog(const Schema& s) {
  auto out = std::make_shared<XProgHost>();
  XProgHost& x = *out;
  Builder b{x, {}, {}};
  x.xbytes.push_back(0);
  const size_t NS = s.scopes.size();
  x.scopes.resize(NS);
  for (size_t i = 0; i < NS; ++i) {
    GkXScope& sc = x.scopes[i];
    sc.gen = 0;
    sc.parent = i ? s.scopes[i].parent : -1;
    sc.first_child = sc.next_sibling = GK_NONE;
    sc.first_col = sc.ncols = 0;
  }
  for (size_t i = 1; i < NS; ++i) {
    const XK k = closure_xinfo(*s.scopes[i].gen).k;
    if (k != XK::Path && k != XK::Elem) throw RegoError{"rego_unsupported: device-ingest: scope generator " + s.scopes[i].gen->key};
    x.scopes[i].gen = b.add(s.scopes[i].gen);
  }
  // child lists in scope order (the order the host flattener fills them in)
  for (size_t i = NS; i-- > 1;) {
    GkXScope& p = x.scopes[s.scopes[i].parent];
    x.scopes[i].next_sibling = p.first_child;
    p.first_child = (uint32_t)i;
  }
  x.cols.resize(s.cols.size());
  for (size_t i = 0; i < s.cols.size(); ++i) {
    GkXCol& c = x.cols[i];
    c.closure = b.add(s.cols[i].expr);
    c.scope = (uint32_t)s.cols[i].scope;
    c.enc = s.cols[i].enc;
    c.bytes_slot = GK_NONE;
    const uint32_t kind = x.cl[c.closure].kind;
    if (c.enc & GK_ENC_BYTES) {
      if (kind == GK_X_LUT || kind == GK_X_COUNT)
        throw RegoError{"rego_unsupported: device-ingest: suffix / contains test on a computed string: " + s.cols[i].expr->key};
      c.bytes_slot = x.nbytecols++;
    }
  }
  for (size_t sc = 0; sc < NS; ++sc) {
    x.scopes[sc].first_col = (uint32_t)x.col_order.size();
    for (size_t i = 0; i < s.cols.size(); ++i)
      if ((size_t)s.cols[i].scope == sc) x.col_order.push_back((uint32_t)i);
    x.scopes[sc].ncols = (uint32_t)x.col_order.size() - x.scopes[sc].first_col;
  }
  if (x.xkeys.empty()) x.xkeys.push_back(0);
  if (x.xargs.empty()) x.xargs.push_back(0);
  if (x.col_order.empty()) x.col_order.push_back(0);
  return out;
}

void build_sid_table(const StringTable& st, SidTable& out) {
  std::vector<uint32_t> off;
  std::vector<uint8_t> bytes;
  st.snapshot(off, bytes);
  const uint32_t n = (uint32_t)off.size() - 1;
  uint32_t cap = 1024;
  while (cap < 4u * n) cap <<= 1;
  out.tab.init(cap);
  out.sid_true = out.sid_false = out.sid_null = GK_SID_OTHER;
  for (uint32_t sid = 2; sid < n; ++sid) {   // 0 / 1 are the reserved "undefined" / "other"
    const uint8_t* p = bytes.data() + off[sid];
    const uint32_t len = off[sid + 1] - off[sid];
    if (len == 0) continue;
    switch (p[0]) {
      case 's': out.tab.put(xhash(GK_SEED_STR, p + 1, len - 1), sid); break;
      case 'n': out.tab.put(xhash(GK_SEED_NUM, p + 1, len - 1), sid); break;
      case 't': out.sid_true = sid; break;
      case 'f': out.sid_false = sid; break;
      case 'z': out.sid_null = sid; break;
      default: break;   // composites: the device answers GK_SID_OTHER for them (documented limit)
    }
  }
  out.nstrings = n;
}

}  // namespace gk
