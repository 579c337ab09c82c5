#include "val.hpp"

#include <mutex>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace gk {

// ------------------------------------------------------------------------------------------- numbers
Num Num::of_double(double v) {
  Num n;
  // Integral doubles below 1e21 are integers to Rego: Go's JSON encoder writes them without an exponent, and OPA keeps
  // the literal.  From 1e21 on the literal has an exponent and the value prints as a float (`1e+30`).
  if (std::isfinite(v) && v == std::floor(v) && std::fabs(v) < 1e21) {
    n.is_int = true;
    n.i = (__int128)v;
    n.d = v;
  } else {
    n.is_int = false;
    n.d = v;
  }
  return n;
}

int num_cmp(const Num& a, const Num& b) {
  if (a.is_int && b.is_int) return a.i < b.i ? -1 : (a.i > b.i ? 1 : 0);
  double x = a.as_double(), y = b.as_double();
  return x < y ? -1 : (x > y ? 1 : 0);
}

static std::string i128_str(__int128 v) {
  if (v == 0) return "0";
  bool neg = v < 0;
  unsigned __int128 u = neg ? (unsigned __int128)(-(v + 1)) + 1 : (unsigned __int128)v;
  char buf[48];
  int p = 47;
  buf[p] = 0;
  while (u) {
    buf[--p] = char('0' + (int)(u % 10));
    u /= 10;
  }
  if (neg) buf[--p] = '-';
  return std::string(buf + p);
}

std::string num_str(const Num& n) {
  if (n.is_int) return i128_str(n.i);
  char buf[64];
  // shortest round-trip repr
  for (int prec = 1; prec <= 17; ++prec) {
    snprintf(buf, sizeof buf, "%.*g", prec, n.d);
    if (strtod(buf, nullptr) == n.d) break;
  }
  return buf;
}

bool num_fits_i64(const Num& n, int64_t* out) {
  if (!n.is_int) return false;
  if (n.i < (__int128)INT64_MIN || n.i > (__int128)INT64_MAX) return false;
  *out = (int64_t)n.i;
  return true;
}

int64_t num_key(const Num& n) {
  const __int128 lim = ((__int128)1 << 62) - 1;
  __int128 fl;
  bool frac = false;
  if (n.is_int) {
    fl = n.i;
  } else if (!(std::fabs(n.d) < 9.0e18)) {   // (also NaN / infinities)
    fl = n.d > 0 ? lim + 1 : -lim - 1;
  } else {
    const double f = std::floor(n.d);
    fl = (__int128)f;
    frac = f != n.d;
  }
  if (fl > lim) return INT64_MAX - 1;
  if (fl < -lim) return INT64_MIN + 1;
  return (int64_t)(2 * fl + (frac ? 1 : 0));
}

// ------------------------------------------------------------------------------------------ node pool
namespace {
// Nodes outlive thread-local storage at process exit (engines are torn down after the main thread's thread_locals): once
// the pool is gone, blocks are simply not recycled.  (A trivially destructible flag stays readable after destruction.)
thread_local bool t_pool_alive = true;
struct BlockPool {
  static std::mutex& mu() { static std::mutex m; return m; }
  static std::vector<void*>& global() { static std::vector<void*> g; return g; }
  std::vector<void*> free_;
  size_t block = 0;
  ~BlockPool() {
    t_pool_alive = false;
    std::lock_guard<std::mutex> l(mu());
    auto& g = global();
    g.insert(g.end(), free_.begin(), free_.end());
  }
  void* get(size_t bytes) {
    if (!block) block = (bytes + 15) / 16 * 16;
    if (bytes > block) return ::operator new(bytes);      // (never: one block type per pool)
    if (free_.empty()) {
      {
        std::lock_guard<std::mutex> l(mu());
        auto& g = global();
        size_t take = std::min<size_t>(g.size(), 1024);
        free_.assign(g.end() - take, g.end());
        g.resize(g.size() - take);
      }
      if (free_.empty()) {
        const size_t n = 512;
        char* slab = static_cast<char*>(::operator new(n * block));
        for (size_t i = 0; i < n; ++i) free_.push_back(slab + i * block);
      }
    }
    void* p = free_.back();
    free_.pop_back();
    return p;
  }
  void put(void* p) { free_.push_back(p); }
};
thread_local BlockPool t_pool;

template <class T>
struct PoolAlloc {
  using value_type = T;
  PoolAlloc() = default;
  template <class U>
  PoolAlloc(const PoolAlloc<U>&) {}
  T* allocate(size_t n) {
    if (n != 1) return static_cast<T*>(::operator new(n * sizeof(T)));
    if (!t_pool_alive) return static_cast<T*>(::operator new((sizeof(T) + 15) / 16 * 16));
    return static_cast<T*>(t_pool.get(sizeof(T)));
  }
  void deallocate(T* p, size_t n) {
    if (n != 1) ::operator delete(p);
    else if (t_pool_alive) t_pool.put(p);
    // else: the pool of this thread is gone (process exit); the block is part of a slab and is left alone
  }
  template <class U>
  bool operator==(const PoolAlloc<U>&) const { return true; }
  template <class U>
  bool operator!=(const PoolAlloc<U>&) const { return false; }
};
}  // namespace

std::shared_ptr<Node> new_node() {
  static const bool pooled = [] {
    const char* e = getenv("GK_NODE_POOL");
    return !(e && e[0] == '0');
  }();
  return pooled ? std::allocate_shared<Node>(PoolAlloc<Node>()) : std::make_shared<Node>();
}

// -------------------------------------------------------------------------------------- constructors
VP v_null() {
  static thread_local VP v = [] { auto n = new_node(); n->t = VT::Null; return VP(n); }();
  return v;
}
VP v_bool(bool b) {
  static thread_local VP t = [] { auto n = new_node(); n->t = VT::True; return VP(n); }();
  static thread_local VP f = [] { auto n = new_node(); n->t = VT::False; return VP(n); }();
  return b ? t : f;
}
VP v_num(const Num& x) {
  auto n = new_node();
  n->t = VT::Num;
  n->n = x;
  return n;
}
VP v_int(long long i) { return v_num(Num::of_int(i)); }
VP v_str(std::string s) {
  auto n = new_node();
  n->t = VT::Str;
  n->s = std::move(s);
  return n;
}
VP v_arr(std::vector<VP> items) {
  auto n = new_node();
  n->t = VT::Arr;
  n->items = std::move(items);
  return n;
}
VP v_set(std::vector<VP> items) {
  std::stable_sort(items.begin(), items.end(), [](const VP& a, const VP& b) { return v_cmp(a, b) < 0; });
  items.erase(std::unique(items.begin(), items.end(), [](const VP& a, const VP& b) { return v_eq(a, b); }),
              items.end());
  auto n = new_node();
  n->t = VT::Set;
  n->items = std::move(items);
  return n;
}
VP v_obj(std::vector<std::pair<VP, VP>> kv) {
  bool sorted = true;
  for (size_t i = 1; i < kv.size() && sorted; ++i) sorted = v_cmp(kv[i - 1].first, kv[i].first) < 0;
  if (sorted) {   // strictly ascending: nothing to sort, nothing to merge
    auto n = new_node();
    n->t = VT::Obj;
    n->kv = std::move(kv);
    return n;
  }
  std::stable_sort(kv.begin(), kv.end(), [](const auto& a, const auto& b) { return v_cmp(a.first, b.first) < 0; });
  // later duplicates win (json.Unmarshal into map semantics)
  std::vector<std::pair<VP, VP>> out;
  for (auto& e : kv) {
    if (!out.empty() && v_eq(out.back().first, e.first)) out.back().second = e.second;
    else out.push_back(std::move(e));
  }
  auto n = new_node();
  n->t = VT::Obj;
  n->kv = std::move(out);
  return n;
}

// ------------------------------------------------------------------------------------------ ordering
int type_rank(VT t) {
  switch (t) {
    case VT::Null: return 0;
    case VT::False:
    case VT::True: return 1;
    case VT::Num: return 2;
    case VT::Str: return 3;
    case VT::Arr: return 4;
    case VT::Obj: return 5;
    case VT::Set: return 6;
    default: return -1;
  }
}

int v_cmp(const VP& a, const VP& b) {
  if (a.get() == b.get()) return 0;
  int ra = type_rank(a->t), rb = type_rank(b->t);
  if (ra != rb) return ra < rb ? -1 : 1;
  switch (a->t) {
    case VT::Null: return 0;
    case VT::False:
    case VT::True: return (int)a->t - (int)b->t;
    case VT::Num: return num_cmp(a->n, b->n);
    case VT::Str: {
      int c = a->s.compare(b->s);
      return c < 0 ? -1 : (c > 0 ? 1 : 0);
    }
    case VT::Arr:
    case VT::Set: {
      size_t n = std::min(a->items.size(), b->items.size());
      for (size_t i = 0; i < n; ++i) {
        int c = v_cmp(a->items[i], b->items[i]);
        if (c) return c;
      }
      return a->items.size() < b->items.size() ? -1 : (a->items.size() > b->items.size() ? 1 : 0);
    }
    case VT::Obj: {
      size_t n = std::min(a->kv.size(), b->kv.size());
      for (size_t i = 0; i < n; ++i) {
        int c = v_cmp(a->kv[i].first, b->kv[i].first);
        if (c) return c;
        c = v_cmp(a->kv[i].second, b->kv[i].second);
        if (c) return c;
      }
      return a->kv.size() < b->kv.size() ? -1 : (a->kv.size() > b->kv.size() ? 1 : 0);
    }
    default: return 0;
  }
}

VP v_deep_copy(const VP& v) {
  if (!v) return v;
  switch (v->t) {
    case VT::Null: return v_null();
    case VT::True: return v_bool(true);
    case VT::False: return v_bool(false);
    case VT::Num: return v_num(v->n);
    case VT::Str: return v_str(v->s);
    case VT::Arr:
    case VT::Set: {
      auto n = new_node();
      n->t = v->t;
      for (auto& x : v->items) n->items.push_back(v_deep_copy(x));
      return n;
    }
    default: {
      auto n = new_node();
      n->t = VT::Obj;
      for (auto& e : v->kv) n->kv.emplace_back(v_deep_copy(e.first), v_deep_copy(e.second));
      return n;
    }
  }
}

VP obj_get(const VP& o, const VP& key) {
  if (!o || o->t != VT::Obj) return nullptr;
  size_t lo = 0, hi = o->kv.size();
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    int c = v_cmp(o->kv[mid].first, key);
    if (c == 0) return o->kv[mid].second;
    if (c < 0) lo = mid + 1;
    else hi = mid;
  }
  return nullptr;
}

VP obj_get(const VP& o, const char* key) {
  if (!o || o->t != VT::Obj) return nullptr;
  size_t lo = 0, hi = o->kv.size();
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    const VP& k = o->kv[mid].first;
    int c;
    if (k->t != VT::Str) c = type_rank(k->t) < 3 ? -1 : 1;
    else {
      c = k->s.compare(key);
    }
    if (c == 0) return o->kv[mid].second;
    if (c < 0) lo = mid + 1;
    else hi = mid;
  }
  return nullptr;
}

VP set_find(const VP& s, const VP& x) {
  if (!s || s->t != VT::Set) return nullptr;
  size_t lo = 0, hi = s->items.size();
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    int c = v_cmp(s->items[mid], x);
    if (c == 0) return s->items[mid];
    if (c < 0) lo = mid + 1;
    else hi = mid;
  }
  return nullptr;
}

// ---------------------------------------------------------------------------------------------- JSON
namespace {
// Per-thread recycling of the nodes a Kubernetes document repeats endlessly (object keys, short enum-like strings,
// small integers): direct-mapped, so a hit costs one hash + one compare and no allocation.
struct StrCache {
  static const size_t N = 2048;
  VP slot[N];
  const VP& get(const char* s, size_t n) {
    // eight bytes per multiply (keys and short values are 2 - 20 bytes: a byte-at-a-time hash was ~10 % of the parse)
    uint64_t h = (uint64_t)n * 0x9E3779B97F4A7C15ull;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
      uint64_t w;
      memcpy(&w, s + i, 8);
      h = (h ^ w) * 0xff51afd7ed558ccdull;
      h ^= h >> 32;
    }
    if (i < n) {
      uint64_t w = 0;
      memcpy(&w, s + i, n - i);
      h = (h ^ w) * 0xff51afd7ed558ccdull;
      h ^= h >> 32;
    }
    VP& v = slot[(h ^ (h >> 29)) & (N - 1)];
    if (!v || v->s.size() != n || memcmp(v->s.data(), s, n) != 0) v = v_str(std::string(s, n));
    return v;
  }
};
struct JP {
  const char* p;
  const char* e;
  int depth = 0;
  std::vector<std::pair<VP, VP>>& kvs;   // scratch stacks: children are collected here, then moved into an exact-size vector
  std::vector<VP>& its;
  StrCache& keys;
  StrCache& shorts;
  [[noreturn]] void fail(const char* m) { throw JsonError{std::string("invalid JSON: ") + m}; }
  // a string without escapes is viewed in place
  bool plain_str(const char*& s, size_t& n) {
    const char* q = p + 1;
    while (q < e && *q != '"' && *q != '\\') ++q;
    if (q >= e || *q != '"') return false;
    s = p + 1;
    n = q - s;
    p = q + 1;
    return true;
  }
  VP str_node(StrCache& cache, size_t max_cached) {
    const char* s;
    size_t n;
    if (plain_str(s, n)) return n <= max_cached ? cache.get(s, n) : v_str(std::string(s, n));
    return v_str(str());
  }
  VP make_obj(size_t base) {
    size_t n = kvs.size() - base;
    auto* a = kvs.data() + base;
    bool sorted = true;
    for (size_t i = 1; i < n && sorted; ++i) sorted = a[i - 1].first->s < a[i].first->s;   // strictly ascending: no duplicates
    if (!sorted) {
      // stable insertion sort on the (small) child list, then later duplicates win (json.Unmarshal into a map)
      for (size_t i = 1; i < n; ++i) {
        auto x = std::move(a[i]);
        size_t j = i;
        while (j > 0 && x.first->s < a[j - 1].first->s) {
          a[j] = std::move(a[j - 1]);
          --j;
        }
        a[j] = std::move(x);
      }
      size_t w = 0;
      for (size_t i = 0; i < n; ++i) {
        if (w > 0 && a[w - 1].first->s == a[i].first->s) a[w - 1].second = std::move(a[i].second);
        else {
          if (w != i) a[w] = std::move(a[i]);
          ++w;
        }
      }
      n = w;
    }
    auto node = new_node();
    node->t = VT::Obj;
    node->kv.assign(std::make_move_iterator(a), std::make_move_iterator(a + n));
    kvs.resize(base);
    return node;
  }
  void ws() {
    while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
  }
  static void utf8(std::string& out, unsigned cp) {
    if (cp < 0x80) out.push_back((char)cp);
    else if (cp < 0x800) {
      out.push_back((char)(0xC0 | (cp >> 6)));
      out.push_back((char)(0x80 | (cp & 0x3F)));
    } else if (cp < 0x10000) {
      out.push_back((char)(0xE0 | (cp >> 12)));
      out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      out.push_back((char)(0x80 | (cp & 0x3F)));
    } else {
      out.push_back((char)(0xF0 | (cp >> 18)));
      out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
      out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      out.push_back((char)(0x80 | (cp & 0x3F)));
    }
  }
  unsigned hex4() {
    if (e - p < 4) fail("bad \\u escape");
    unsigned v = 0;
    for (int i = 0; i < 4; ++i) {
      char c = *p++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else fail("bad \\u escape");
    }
    return v;
  }
  std::string str() {
    // *p == '"'
    ++p;
    std::string out;
    const char* s = p;
    while (p < e && *p != '"' && *p != '\\') ++p;
    out.assign(s, p - s);
    while (p < e && *p != '"') {
      if (*p == '\\') {
        ++p;
        if (p >= e) fail("bad escape");
        char c = *p++;
        switch (c) {
          case 'n': out.push_back('\n'); break;
          case 't': out.push_back('\t'); break;
          case 'r': out.push_back('\r'); break;
          case 'b': out.push_back('\b'); break;
          case 'f': out.push_back('\f'); break;
          case '/': out.push_back('/'); break;
          case '\\': out.push_back('\\'); break;
          case '"': out.push_back('"'); break;
          case 'u': {
            unsigned cp = hex4();
            if (cp >= 0xD800 && cp < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
              p += 2;
              unsigned lo = hex4();
              if (lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
              else cp = 0xFFFD;
            }
            utf8(out, cp);
            break;
          }
          default: fail("bad escape");
        }
      } else {
        out.push_back(*p++);
      }
    }
    if (p >= e) fail("unterminated string");
    ++p;
    return out;
  }
  VP value() {
    ws();
    if (p >= e) fail("unexpected end");
    if (++depth > 512) fail("too deep");
    VP r;
    char c = *p;
    if (c == '{') {
      ++p;
      const size_t base = kvs.size();
      ws();
      if (p < e && *p == '}') {
        ++p;
      } else {
        while (true) {
          ws();
          if (p >= e || *p != '"') fail("object key expected");
          VP k = str_node(keys, 48);
          ws();
          if (p >= e || *p != ':') fail("':' expected");
          ++p;
          VP v = value();
          kvs.emplace_back(std::move(k), std::move(v));
          ws();
          if (p < e && *p == ',') {
            ++p;
            continue;
          }
          if (p < e && *p == '}') {
            ++p;
            break;
          }
          fail("',' or '}' expected");
        }
      }
      r = make_obj(base);
    } else if (c == '[') {
      ++p;
      const size_t base = its.size();
      ws();
      if (p < e && *p == ']') {
        ++p;
      } else {
        while (true) {
          VP x = value();
          its.push_back(std::move(x));
          ws();
          if (p < e && *p == ',') {
            ++p;
            continue;
          }
          if (p < e && *p == ']') {
            ++p;
            break;
          }
          fail("',' or ']' expected");
        }
      }
      {
        auto node = new_node();
        node->t = VT::Arr;
        node->items.assign(std::make_move_iterator(its.begin() + base), std::make_move_iterator(its.end()));
        its.resize(base);
        r = std::move(node);
      }
    } else if (c == '"') {
      r = str_node(shorts, 12);
    } else if (c == 't' && e - p >= 4 && !memcmp(p, "true", 4)) {
      p += 4;
      r = v_bool(true);
    } else if (c == 'f' && e - p >= 5 && !memcmp(p, "false", 5)) {
      p += 5;
      r = v_bool(false);
    } else if (c == 'n' && e - p >= 4 && !memcmp(p, "null", 4)) {
      p += 4;
      r = v_null();
    } else if (c == '-' || (c >= '0' && c <= '9')) {
      const char* s = p;
      bool isint = true;
      if (*p == '-') ++p;
      while (p < e && *p >= '0' && *p <= '9') ++p;
      if (p < e && (*p == '.' || *p == 'e' || *p == 'E')) {
        isint = false;
        while (p < e && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) ++p;
      }
      size_t len = p - s;
      if (isint && len <= 37) {
        __int128 v = 0;
        const char* q = s;
        bool neg = *q == '-';
        if (neg) ++q;
        if (q == p) fail("bad number");
        for (; q < p; ++q) v = v * 10 + (*q - '0');
        r = v_num(Num::of_int(neg ? -v : v));
      } else {
        std::string tmp(s, len);
        char* endp = nullptr;
        double d = strtod(tmp.c_str(), &endp);
        if (endp != tmp.c_str() + len) fail("bad number");
        r = v_num(Num::of_double(d));
      }
    } else {
      fail("unexpected character");
    }
    --depth;
    return r;
  }
};
}  // namespace

VP json_parse(const char* p, size_t n) {
  static thread_local std::vector<std::pair<VP, VP>> kvs;
  static thread_local std::vector<VP> its;
  static thread_local StrCache keys, shorts;
  kvs.clear();   // (a failed parse leaves its partial children behind)
  its.clear();
  JP jp{p, p + n, 0, kvs, its, keys, shorts};
  VP v = jp.value();
  jp.ws();
  if (jp.p != jp.e) jp.fail("trailing characters");
  return v;
}

void json_quote(const std::string& s, std::string& out) {
  out.push_back('"');
  for (unsigned char c : s) {
    switch (c) {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      default:
        if (c < 0x20) {
          char b[8];
          snprintf(b, sizeof b, "\\u%04x", c);
          out += b;
        } else {
          out.push_back((char)c);
        }
    }
  }
  out.push_back('"');
}

static void json_rec(const VP& v, std::string& out) {
  switch (v->t) {
    case VT::Null: out += "null"; break;
    case VT::True: out += "true"; break;
    case VT::False: out += "false"; break;
    case VT::Num: out += num_str(v->n); break;
    case VT::Str: json_quote(v->s, out); break;
    case VT::Arr:
    case VT::Set: {
      out.push_back('[');
      bool first = true;
      for (auto& x : v->items) {
        if (!first) out.push_back(',');
        first = false;
        json_rec(x, out);
      }
      out.push_back(']');
      break;
    }
    case VT::Obj: {
      out.push_back('{');
      bool first = true;
      for (auto& e : v->kv) {
        if (!first) out.push_back(',');
        first = false;
        if (e.first->t == VT::Str) json_quote(e.first->s, out);
        else json_quote(fmt_value(e.first, false), out);
        out.push_back(':');
        json_rec(e.second, out);
      }
      out.push_back('}');
      break;
    }
    default: out += "null";
  }
}

std::string json_str(const VP& v) {
  std::string out;
  if (!v) return "null";
  json_rec(v, out);
  return out;
}

// ------------------------------------------------------------------------------------------ `%v`
static void fmt_rec(const VP& v, bool top, std::string& out) {
  switch (v->t) {
    case VT::Null: out += "null"; break;
    case VT::True: out += "true"; break;
    case VT::False: out += "false"; break;
    case VT::Num: out += num_str(v->n); break;
    case VT::Str:
      if (top) out += v->s;
      else json_quote(v->s, out);
      break;
    case VT::Arr: {
      out.push_back('[');
      for (size_t i = 0; i < v->items.size(); ++i) {
        if (i) out += ", ";
        fmt_rec(v->items[i], false, out);
      }
      out.push_back(']');
      break;
    }
    case VT::Set: {
      if (v->items.empty()) {
        out += "set()";
        break;
      }
      out.push_back('{');
      for (size_t i = 0; i < v->items.size(); ++i) {
        if (i) out += ", ";
        fmt_rec(v->items[i], false, out);
      }
      out.push_back('}');
      break;
    }
    case VT::Obj: {
      out.push_back('{');
      for (size_t i = 0; i < v->kv.size(); ++i) {
        if (i) out += ", ";
        fmt_rec(v->kv[i].first, false, out);
        out += ": ";
        fmt_rec(v->kv[i].second, false, out);
      }
      out.push_back('}');
      break;
    }
    default: break;
  }
}

std::string fmt_value(const VP& v, bool top) {
  std::string out;
  if (v) fmt_rec(v, top, out);
  return out;
}

std::string intern_key(const VP& v) {
  switch (v->t) {
    case VT::Null: return "z";
    case VT::True: return "t";
    case VT::False: return "f";
    case VT::Num: return "n" + num_str(v->n);
    case VT::Str: return "s" + v->s;
    default: return "j" + fmt_value(v, false);
  }
}

}  // namespace gk
