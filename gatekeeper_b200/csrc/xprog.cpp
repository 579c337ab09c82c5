// Extraction program builder: compiled Schema (closures classified by lower.cpp's closure_xinfo) -> the flat tables the
// ingest kernels interpret (ingest_core.h GkXProg).
#include "xprog.hpp"

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
  uint32_t add_path(uint32_t base, const std::vector<VP>& keys) {
    if (keys.empty()) return base;
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
  Builder b{x, {}};
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
