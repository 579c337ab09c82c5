// Concrete top-down evaluation of the Rego subset + builtins.  See rego.hpp for where this is (and is
// not) used.  Semantics restate OPA v1.13.2's documented behaviour: undefined propagation, negation as
// failure, partial-set extents, multi-definition functions as OR, type errors => undefined.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <regex>
#include <unordered_map>

#include "rego.hpp"

namespace gk {

// ------------------------------------------------------------------------------------------ builtins
namespace {

inline bool is_num(const VP& v) { return v->t == VT::Num; }
inline bool is_str(const VP& v) { return v->t == VT::Str; }
inline bool is_coll(const VP& v) { return v->t == VT::Arr || v->t == VT::Set; }

VP arith(const std::string& op, const VP& a, const VP& b) {
  if (a->t == VT::Set && b->t == VT::Set && (op == "minus" || op == "and" || op == "or")) {
    std::vector<VP> out;
    if (op == "minus") {
      for (auto& x : a->items)
        if (!set_find(b, x)) out.push_back(x);
    } else if (op == "and") {
      for (auto& x : a->items)
        if (set_find(b, x)) out.push_back(x);
    } else {
      out = a->items;
      out.insert(out.end(), b->items.begin(), b->items.end());
    }
    return v_set(std::move(out));
  }
  if (!is_num(a) || !is_num(b)) return nullptr;
  const Num &x = a->n, &y = b->n;
  if (x.is_int && y.is_int) {
    __int128 r;
    if (op == "plus") {
      if (__builtin_add_overflow(x.i, y.i, &r)) return v_num(Num::of_double(x.d + y.d));
      return v_num(Num::of_int(r));
    }
    if (op == "minus") {
      if (__builtin_sub_overflow(x.i, y.i, &r)) return v_num(Num::of_double(x.d - y.d));
      return v_num(Num::of_int(r));
    }
    if (op == "mul") {
      if (__builtin_mul_overflow(x.i, y.i, &r)) return v_num(Num::of_double(x.d * y.d));
      return v_num(Num::of_int(r));
    }
    if (op == "div") {
      if (y.i == 0) return nullptr;
      if (x.i % y.i == 0) return v_num(Num::of_int(x.i / y.i));
      return v_num(Num::of_double((double)x.i / (double)y.i));
    }
    if (op == "rem") {
      if (y.i == 0) return nullptr;
      return v_num(Num::of_int(x.i % y.i));   // C++ % truncates toward zero like Go's big.Int.Rem
    }
    return nullptr;
  }
  double p = x.as_double(), q = y.as_double();
  if (op == "plus") return v_num(Num::of_double(p + q));
  if (op == "minus") return v_num(Num::of_double(p - q));
  if (op == "mul") return v_num(Num::of_double(p * q));
  if (op == "div") return q == 0 ? nullptr : v_num(Num::of_double(p / q));
  return nullptr;
}

size_t utf8_len(const std::string& s) {
  size_t n = 0;
  for (unsigned char c : s)
    if ((c & 0xC0) != 0x80) ++n;
  return n;
}
// byte offset of the cp-th code point
size_t utf8_off(const std::string& s, size_t cp) {
  size_t i = 0, n = 0;
  while (i < s.size() && n < cp) {
    ++i;
    while (i < s.size() && ((unsigned char)s[i] & 0xC0) == 0x80) ++i;
    ++n;
  }
  return i;
}

VP to_number(const VP& x) {
  switch (x->t) {
    case VT::Null: return v_int(0);
    case VT::False: return v_int(0);
    case VT::True: return v_int(1);
    case VT::Num: return x;
    case VT::Str: {
      const std::string& s = x->s;
      if (s.empty()) return nullptr;
      size_t i = 0;
      if (s[i] == '+' || s[i] == '-') ++i;
      bool digits = false, dot = false, exp = false, ok = true;
      for (; i < s.size(); ++i) {
        char c = s[i];
        if (c >= '0' && c <= '9') digits = true;
        else if (c == '.' && !dot && !exp) dot = true;
        else if ((c == 'e' || c == 'E') && digits && !exp) {
          exp = true;
          if (i + 1 < s.size() && (s[i + 1] == '+' || s[i + 1] == '-')) ++i;
          if (i + 1 >= s.size()) ok = false;
        } else {
          ok = false;
          break;
        }
      }
      if (!ok || !digits) return nullptr;
      if (!dot && !exp) {
        std::string t = s[0] == '+' ? s.substr(1) : s;
        try {
          return json_parse(t.data(), t.size());
        } catch (JsonError&) {
          // leading zeros etc.: fall through to strtod
        }
      }
      return v_num(Num::of_double(strtod(s.c_str(), nullptr)));
    }
    default: return nullptr;
  }
}

std::string trim_cutset(const std::string& s, const std::string& cut, bool left, bool right) {
  size_t b = 0, e = s.size();
  if (left)
    while (b < e && cut.find(s[b]) != std::string::npos) ++b;
  if (right)
    while (e > b && cut.find(s[e - 1]) != std::string::npos) --e;
  return s.substr(b, e - b);
}

std::string replace_all(const std::string& s, const std::string& from, const std::string& to) {
  if (from.empty()) {
    // Go strings.ReplaceAll with empty `old` inserts `new` between every rune and at both ends
    std::string out = to;
    size_t i = 0;
    while (i < s.size()) {
      size_t j = i + 1;
      while (j < s.size() && ((unsigned char)s[j] & 0xC0) == 0x80) ++j;
      out.append(s, i, j - i);
      out += to;
      i = j;
    }
    return out;
  }
  std::string out;
  size_t i = 0;
  while (true) {
    size_t j = s.find(from, i);
    if (j == std::string::npos) break;
    out.append(s, i, j - i);
    out += to;
    i = j + from.size();
  }
  out.append(s, i, std::string::npos);
  return out;
}

VP split(const std::string& s, const std::string& d) {
  std::vector<VP> out;
  if (d.empty()) {
    size_t i = 0;
    while (i < s.size()) {
      size_t j = i + 1;
      while (j < s.size() && ((unsigned char)s[j] & 0xC0) == 0x80) ++j;
      out.push_back(v_str(s.substr(i, j - i)));
      i = j;
    }
    return v_arr(std::move(out));
  }
  size_t i = 0;
  while (true) {
    size_t j = s.find(d, i);
    if (j == std::string::npos) break;
    out.push_back(v_str(s.substr(i, j - i)));
    i = j + d.size();
  }
  out.push_back(v_str(s.substr(i)));
  return v_arr(std::move(out));
}

// OPA sprintf == Go fmt.Sprintf over ast.Values: only the verbs the in-tree templates use.
VP sprintf_(const VP& f, const VP& args) {
  if (!is_str(f) || args->t != VT::Arr) return nullptr;
  const std::string& fmt = f->s;
  std::string out;
  size_t ai = 0;
  auto go_type = [](const VP& a) { return a->t == VT::Str ? "string" : "ast.Value"; };
  for (size_t i = 0; i < fmt.size(); ++i) {
    char c = fmt[i];
    if (c != '%') {
      out.push_back(c);
      continue;
    }
    if (++i >= fmt.size()) {
      out += "%!(NOVERB)";
      break;
    }
    char v = fmt[i];
    if (v == '%') {
      out.push_back('%');
      continue;
    }
    if (ai >= args->items.size()) {
      out += "%!";
      out.push_back(v);
      out += "(MISSING)";
      continue;
    }
    const VP& a = args->items[ai++];
    if (v == 'v' || v == 's') out += fmt_value(a, true);
    else if (v == 'd') {
      if (a->t == VT::Num && a->n.is_int) out += num_str(a->n);
      else out += std::string("%!d(") + go_type(a) + "=" + fmt_value(a, true) + ")";
    } else if (v == 'q' && a->t == VT::Str) json_quote(a->s, out);
    else throw RegoError{std::string("rego_unsupported: sprintf verb %") + v};
  }
  if (ai < args->items.size()) {
    out += "%!(EXTRA ";
    for (size_t j = ai; j < args->items.size(); ++j) {
      if (j > ai) out += ", ";
      out += std::string(go_type(args->items[j])) + "=" + fmt_value(args->items[j], true);
    }
    out += ")";
  }
  return v_str(out);
}

bool str_list(const VP& v, std::vector<const std::string*>& out) {
  if (is_str(v)) {
    out.push_back(&v->s);
    return true;
  }
  if (!is_coll(v)) return false;
  for (auto& x : v->items) {
    if (!is_str(x)) return false;
    out.push_back(&x->s);
  }
  return true;
}

bool starts_with(const std::string& s, const std::string& p) { return s.size() >= p.size() && !s.compare(0, p.size(), p); }
bool ends_with(const std::string& s, const std::string& p) {
  return s.size() >= p.size() && !s.compare(s.size() - p.size(), p.size(), p);
}

VP re_match(const VP& pat, const VP& s) {
  if (!is_str(pat) || !is_str(s)) return nullptr;
  // the all-digits patterns of the quantity parsers need no regex engine at all
  if (pat->s == "^[0-9]+$" || pat->s == "^\\d+$") {
    if (s->s.empty()) return v_bool(false);
    for (char c : s->s)
      if (c < '0' || c > '9') return v_bool(false);
    return v_bool(true);
  }
  try {
    // RE2 syntax is close to ECMAScript for the anchors/classes/quantifiers the fixtures use.
    // Compiled patterns are cached per thread: constructing a std::regex costs tens of microseconds.
    static thread_local std::unordered_map<std::string, std::regex> cache;
    auto it = cache.find(pat->s);
    if (it == cache.end()) {
      if (cache.size() > 256) cache.clear();
      it = cache.emplace(pat->s, std::regex(pat->s, std::regex::ECMAScript)).first;
    }
    return v_bool(std::regex_search(s->s, it->second));
  } catch (std::regex_error&) {
    return nullptr;
  }
}

VP object_get(const VP& o, const VP& k, const VP& d) {
  if (o->t != VT::Obj) return nullptr;
  if (k->t == VT::Arr) {
    VP cur = o;
    for (auto& p : k->items) {
      VP nx;
      if (cur->t == VT::Obj) nx = obj_get(cur, p);
      else if (cur->t == VT::Arr && p->t == VT::Num) {
        int64_t ix;
        if (num_fits_i64(p->n, &ix) && ix >= 0 && (size_t)ix < cur->items.size()) nx = cur->items[ix];
      }
      if (!nx) return d;
      cur = nx;
    }
    return cur;
  }
  VP r = obj_get(o, k);
  return r ? r : d;
}

}  // namespace

bool is_builtin(const std::string& name) {
  bool known = true;
  (void)call_builtin(name, {}, &known);
  return known;
}

// ---- collection / object / rounding builtins of OPA's policy reference (v1.x) used by the gatekeeper policy library
static bool int_arg(const VP& v, int64_t* out) { return v->t == VT::Num && num_fits_i64(v->n, out); }
static VP object_union(const VP& a, const VP& b) {
  std::vector<std::pair<VP, VP>> kv(a->kv.begin(), a->kv.end());
  for (auto& e : b->kv) {
    bool merged = false;
    for (auto& x : kv)
      if (v_eq(x.first, e.first)) {
        x.second = (x.second->t == VT::Obj && e.second->t == VT::Obj) ? object_union(x.second, e.second) : e.second;
        merged = true;
        break;
      }
    if (!merged) kv.emplace_back(e.first, e.second);
  }
  return v_obj(std::move(kv));
}
static bool key_list(const VP& ks, std::vector<VP>& out) {
  if (ks->t == VT::Obj) {
    for (auto& e : ks->kv) out.push_back(e.first);
    return true;
  }
  if (!is_coll(ks)) return false;
  out = ks->items;
  return true;
}
static VP object_pick(const VP& o, const VP& ks, bool keep) {
  std::vector<VP> keys;
  if (o->t != VT::Obj || !key_list(ks, keys)) return nullptr;
  std::vector<std::pair<VP, VP>> kv;
  for (auto& e : o->kv) {
    bool has = false;
    for (auto& k : keys) has = has || v_eq(k, e.first);
    if (has == keep) kv.emplace_back(e.first, e.second);
  }
  return v_obj(std::move(kv));
}
static std::string utf8_reverse(const std::string& s) {
  std::string out;
  out.reserve(s.size());
  size_t i = s.size();
  while (i > 0) {
    size_t j = i - 1;
    while (j > 0 && ((unsigned char)s[j] & 0xC0) == 0x80) --j;
    out.append(s, j, i - j);
    i = j;
  }
  return out;
}
static VP round_like(const VP& x, int mode) {   // 0 round (half away from zero, Go math.Round), 1 floor, 2 ceil
  if (!is_num(x)) return nullptr;
  if (x->n.is_int) return x;
  const double d = x->n.d;
  return v_num(Num::of_double(mode == 0 ? std::round(d) : mode == 1 ? std::floor(d) : std::ceil(d)));
}
static VP format_int(const VP& x, const VP& base) {
  int64_t b;
  if (!is_num(x) || !int_arg(base, &b) || (b != 2 && b != 8 && b != 10 && b != 16)) return nullptr;
  __int128 n;
  if (x->n.is_int) n = x->n.i;
  else {
    const double f = std::floor(x->n.d);
    if (!(std::fabs(f) < 1e30)) return nullptr;
    n = (__int128)f;
  }
  const bool neg = n < 0;
  unsigned __int128 u = neg ? (unsigned __int128)(-n) : (unsigned __int128)n;
  std::string out;
  do {
    out.insert(out.begin(), "0123456789abcdef"[(int)(u % (unsigned)b)]);
    u /= (unsigned)b;
  } while (u != 0);
  return v_str((neg ? "-" : "") + out);
}
static const char kB64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
static std::string b64_encode(const std::string& in) {
  std::string out;
  size_t i = 0;
  for (; i + 2 < in.size(); i += 3) {
    const uint32_t v = ((unsigned char)in[i] << 16) | ((unsigned char)in[i + 1] << 8) | (unsigned char)in[i + 2];
    out += kB64[v >> 18], out += kB64[(v >> 12) & 63], out += kB64[(v >> 6) & 63], out += kB64[v & 63];
  }
  if (i + 1 == in.size()) {
    const uint32_t v = (unsigned char)in[i] << 16;
    out += kB64[v >> 18], out += kB64[(v >> 12) & 63], out += "==";
  } else if (i + 2 == in.size()) {
    const uint32_t v = ((unsigned char)in[i] << 16) | ((unsigned char)in[i + 1] << 8);
    out += kB64[v >> 18], out += kB64[(v >> 12) & 63], out += kB64[(v >> 6) & 63], out += '=';
  }
  return out;
}
static bool b64_decode(const std::string& in, std::string& out) {   // StdEncoding: padded, strict alphabet
  if (in.size() % 4) return false;
  auto val = [](char c) -> int {
    const char* p = c ? strchr(kB64, c) : nullptr;
    return p ? (int)(p - kB64) : -1;
  };
  for (size_t i = 0; i < in.size(); i += 4) {
    int v[4], pad = 0;
    for (int k = 0; k < 4; ++k) {
      if (in[i + k] == '=') {
        if (i + 4 != in.size() || k < 2) return false;
        v[k] = 0;
        ++pad;
      } else {
        if (pad) return false;
        v[k] = val(in[i + k]);
        if (v[k] < 0) return false;
      }
    }
    const uint32_t w = (v[0] << 18) | (v[1] << 12) | (v[2] << 6) | v[3];
    out += (char)(w >> 16);
    if (pad < 2) out += (char)((w >> 8) & 255);
    if (pad < 1) out += (char)(w & 255);
  }
  return true;
}

static constexpr uint64_t name_hash(const char* s) {
  uint64_t h = 1469598103934665603ull;
  for (; *s; ++s) h = (h ^ (unsigned char)*s) * 1099511628211ull;
  return h;
}
VP call_builtin(const std::string& n, const std::vector<VP>& a, bool* known) {
  *known = true;
  auto need = [&](size_t k) { return a.size() == k; };
  // names are dispatched on a 64-bit hash (compile-time constants on the right-hand side), confirmed by one compare
  const uint64_t nh = name_hash(n.c_str());
#define IS(name) (nh == std::integral_constant<uint64_t, name_hash(name)>::value && n == name)
#define B(name, arity) if (IS(name)) { if (!need(arity)) return nullptr;
#define E }
  B("equal", 2) return v_bool(v_eq(a[0], a[1])); E
  B("neq", 2) return v_bool(!v_eq(a[0], a[1])); E
  B("lt", 2) return v_bool(v_cmp(a[0], a[1]) < 0); E
  B("lte", 2) return v_bool(v_cmp(a[0], a[1]) <= 0); E
  B("gt", 2) return v_bool(v_cmp(a[0], a[1]) > 0); E
  B("gte", 2) return v_bool(v_cmp(a[0], a[1]) >= 0); E
  if (IS("plus") || IS("minus") || IS("mul") || IS("div") || IS("rem") || IS("and") || IS("or")) {
    if (!need(2)) return nullptr;
    return arith(n, a[0], a[1]);
  }
  B("count", 1)
    switch (a[0]->t) {
      case VT::Str: return v_int((long long)utf8_len(a[0]->s));
      case VT::Arr:
      case VT::Set: return v_int((long long)a[0]->items.size());
      case VT::Obj: return v_int((long long)a[0]->kv.size());
      default: return nullptr;
    }
  E
  B("sprintf", 2) return sprintf_(a[0], a[1]); E
  B("startswith", 2) if (!is_str(a[0]) || !is_str(a[1])) return nullptr; return v_bool(starts_with(a[0]->s, a[1]->s)); E
  B("endswith", 2) if (!is_str(a[0]) || !is_str(a[1])) return nullptr; return v_bool(ends_with(a[0]->s, a[1]->s)); E
  B("contains", 2) if (!is_str(a[0]) || !is_str(a[1])) return nullptr; return v_bool(a[0]->s.find(a[1]->s) != std::string::npos); E
  if (IS("strings.any_prefix_match") || IS("strings.any_suffix_match")) {
    if (!need(2)) return nullptr;
    std::vector<const std::string*> ss, bb;
    if (!str_list(a[0], ss) || !str_list(a[1], bb)) return nullptr;
    bool pre = n == "strings.any_prefix_match";
    for (auto s : ss)
      for (auto b : bb)
        if (pre ? starts_with(*s, *b) : ends_with(*s, *b)) return v_bool(true);
    return v_bool(false);
  }
  B("split", 2) if (!is_str(a[0]) || !is_str(a[1])) return nullptr; return split(a[0]->s, a[1]->s); E
  B("trim", 2) if (!is_str(a[0]) || !is_str(a[1])) return nullptr; return v_str(trim_cutset(a[0]->s, a[1]->s, true, true)); E
  B("trim_left", 2) if (!is_str(a[0]) || !is_str(a[1])) return nullptr; return v_str(trim_cutset(a[0]->s, a[1]->s, true, false)); E
  B("trim_right", 2) if (!is_str(a[0]) || !is_str(a[1])) return nullptr; return v_str(trim_cutset(a[0]->s, a[1]->s, false, true)); E
  B("trim_space", 1) if (!is_str(a[0])) return nullptr; return v_str(trim_cutset(a[0]->s, " \t\n\r\v\f", true, true)); E
  B("trim_prefix", 2)
    if (!is_str(a[0]) || !is_str(a[1])) return nullptr;
    return starts_with(a[0]->s, a[1]->s) ? v_str(a[0]->s.substr(a[1]->s.size())) : a[0];
  E
  B("trim_suffix", 2)
    if (!is_str(a[0]) || !is_str(a[1])) return nullptr;
    return ends_with(a[0]->s, a[1]->s) ? v_str(a[0]->s.substr(0, a[0]->s.size() - a[1]->s.size())) : a[0];
  E
  B("replace", 3)
    if (!is_str(a[0]) || !is_str(a[1]) || !is_str(a[2])) return nullptr;
    return v_str(replace_all(a[0]->s, a[1]->s, a[2]->s));
  E
  B("substring", 3)
    if (!is_str(a[0]) || !is_num(a[1]) || !is_num(a[2])) return nullptr;
    {
      int64_t off, len;
      if (!num_fits_i64(a[1]->n, &off) || !num_fits_i64(a[2]->n, &len) || off < 0) return nullptr;
      size_t total = utf8_len(a[0]->s);
      if ((size_t)off >= total) return v_str("");
      size_t b = utf8_off(a[0]->s, (size_t)off);
      if (len < 0) return v_str(a[0]->s.substr(b));
      size_t e = utf8_off(a[0]->s, std::min(total, (size_t)off + (size_t)len));
      return v_str(a[0]->s.substr(b, e - b));
    }
  E
  B("lower", 1)
    if (!is_str(a[0])) return nullptr;
    { std::string s = a[0]->s; for (auto& c : s) if (c >= 'A' && c <= 'Z') c = char(c + 32); return v_str(s); }
  E
  B("upper", 1)
    if (!is_str(a[0])) return nullptr;
    { std::string s = a[0]->s; for (auto& c : s) if (c >= 'a' && c <= 'z') c = char(c - 32); return v_str(s); }
  E
  B("concat", 2)
    if (!is_str(a[0]) || !is_coll(a[1])) return nullptr;
    {
      std::string out;
      for (size_t i = 0; i < a[1]->items.size(); ++i) {
        if (!is_str(a[1]->items[i])) return nullptr;
        if (i) out += a[0]->s;
        out += a[1]->items[i]->s;
      }
      return v_str(out);
    }
  E
  B("indexof", 2)
    if (!is_str(a[0]) || !is_str(a[1])) return nullptr;
    {
      size_t p = a[0]->s.find(a[1]->s);
      if (p == std::string::npos) return v_int(-1);
      return v_int((long long)utf8_len(a[0]->s.substr(0, p)));
    }
  E
  if (IS("re_match") || IS("regex.match")) {
    if (!need(2)) return nullptr;
    return re_match(a[0], a[1]);
  }
  B("to_number", 1) return to_number(a[0]); E
  B("is_number", 1) return v_bool(a[0]->t == VT::Num); E
  B("is_string", 1) return v_bool(a[0]->t == VT::Str); E
  B("is_boolean", 1) return v_bool(a[0]->t == VT::True || a[0]->t == VT::False); E
  B("is_array", 1) return v_bool(a[0]->t == VT::Arr); E
  B("is_object", 1) return v_bool(a[0]->t == VT::Obj); E
  B("is_set", 1) return v_bool(a[0]->t == VT::Set); E
  B("is_null", 1) return v_bool(a[0]->t == VT::Null); E
  B("any", 1)
    if (!is_coll(a[0])) return nullptr;
    for (auto& x : a[0]->items) if (x->t == VT::True) return v_bool(true);
    return v_bool(false);
  E
  B("all", 1)
    if (!is_coll(a[0])) return nullptr;
    for (auto& x : a[0]->items) if (x->t != VT::True) return v_bool(false);
    return v_bool(true);
  E
  B("object.get", 3) return object_get(a[0], a[1], a[2]); E
  B("array.concat", 2)
    if (a[0]->t != VT::Arr || a[1]->t != VT::Arr) return nullptr;
    { auto v = a[0]->items; v.insert(v.end(), a[1]->items.begin(), a[1]->items.end()); return v_arr(std::move(v)); }
  E
  B("abs", 1)
    if (!is_num(a[0])) return nullptr;
    return a[0]->n.is_int ? v_num(Num::of_int(a[0]->n.i < 0 ? -a[0]->n.i : a[0]->n.i)) : v_num(Num::of_double(std::fabs(a[0]->n.d)));
  E
  if (IS("max") || IS("min")) {
    if (!need(1) || !is_coll(a[0]) || a[0]->items.empty()) return nullptr;
    VP best = a[0]->items[0];
    for (auto& x : a[0]->items)
      if ((n == "max") ? v_cmp(x, best) > 0 : v_cmp(x, best) < 0) best = x;
    return best;
  }
  B("sum", 1)
    if (!is_coll(a[0])) return nullptr;
    { VP acc = v_int(0); for (auto& x : a[0]->items) { acc = arith("plus", acc, x); if (!acc) return nullptr; } return acc; }
  E
  B("internal.member_2", 2)
    if (a[1]->t == VT::Obj) { for (auto& e : a[1]->kv) if (v_eq(e.second, a[0])) return v_bool(true); return v_bool(false); }
    if (is_coll(a[1])) { for (auto& x : a[1]->items) if (v_eq(x, a[0])) return v_bool(true); return v_bool(false); }
    return v_bool(false);
  E
  B("sort", 1)
    if (!is_coll(a[0])) return nullptr;
    { auto v = a[0]->items; std::stable_sort(v.begin(), v.end(), [](const VP& x, const VP& y) { return v_cmp(x, y) < 0; }); return v_arr(std::move(v)); }
  E
  B("object.keys", 1)
    if (a[0]->t != VT::Obj) return nullptr;
    { std::vector<VP> ks; for (auto& e : a[0]->kv) ks.push_back(e.first); return v_set(std::move(ks)); }
  E
  B("object.union", 2) if (a[0]->t != VT::Obj || a[1]->t != VT::Obj) return nullptr; return object_union(a[0], a[1]); E
  B("object.remove", 2) return object_pick(a[0], a[1], false); E
  B("object.filter", 2) return object_pick(a[0], a[1], true); E
  B("numbers.range", 2)
    {
      int64_t lo, hi;
      if (!int_arg(a[0], &lo) || !int_arg(a[1], &hi)) return nullptr;
      if ((lo < hi ? hi - lo : lo - hi) > 1000000) return nullptr;
      std::vector<VP> v;
      if (lo <= hi) for (int64_t i = lo; i <= hi; ++i) v.push_back(v_int(i));
      else for (int64_t i = lo; i >= hi; --i) v.push_back(v_int(i));
      return v_arr(std::move(v));
    }
  E
  B("array.slice", 3)
    {
      int64_t lo, hi;
      if (a[0]->t != VT::Arr || !int_arg(a[1], &lo) || !int_arg(a[2], &hi)) return nullptr;
      lo = std::max<int64_t>(lo, 0);
      hi = std::min<int64_t>(hi, (int64_t)a[0]->items.size());
      if (lo >= hi) return v_arr({});
      return v_arr(std::vector<VP>(a[0]->items.begin() + lo, a[0]->items.begin() + hi));
    }
  E
  B("array.reverse", 1) if (a[0]->t != VT::Arr) return nullptr; return v_arr(std::vector<VP>(a[0]->items.rbegin(), a[0]->items.rend())); E
  B("strings.reverse", 1) if (!is_str(a[0])) return nullptr; return v_str(utf8_reverse(a[0]->s)); E
  B("round", 1) return round_like(a[0], 0); E
  B("floor", 1) return round_like(a[0], 1); E
  B("ceil", 1) return round_like(a[0], 2); E
  B("format_int", 2) return format_int(a[0], a[1]); E
  if (IS("union") || IS("intersection")) {
    if (!need(1) || a[0]->t != VT::Set) return nullptr;
    for (auto& x : a[0]->items) if (x->t != VT::Set) return nullptr;
    if (a[0]->items.empty()) return v_set({});
    VP acc = a[0]->items[0];
    for (size_t i = 1; i < a[0]->items.size(); ++i) acc = arith(n == "union" ? "or" : "and", acc, a[0]->items[i]);
    return acc;
  }
  B("product", 1)
    if (!is_coll(a[0])) return nullptr;
    { VP acc = v_int(1); for (auto& x : a[0]->items) { acc = arith("mul", acc, x); if (!acc) return nullptr; } return acc; }
  E
  B("type_name", 1)
    {
      static const char* names[] = {"", "null", "boolean", "boolean", "number", "string", "array", "object", "set"};
      return v_str(names[(int)a[0]->t]);
    }
  E
  B("base64.encode", 1) if (!is_str(a[0])) return nullptr; return v_str(b64_encode(a[0]->s)); E
  B("base64.decode", 1)
    if (!is_str(a[0])) return nullptr;
    { std::string out; if (!b64_decode(a[0]->s, out)) return nullptr; return v_str(out); }
  E
  if (IS("print") || IS("trace")) return v_bool(true);
#undef B
#undef E
#undef IS
  *known = false;
  return nullptr;
}

// ----------------------------------------------------------------------------------------- evaluator
bool Eval::var_unbound(const Term& t, const Env& env) const {
  if (t.k != TK::Var || env.find(t.vid) || t.vid == m_.vid_input || t.vid == m_.vid_data) return false;
  signed char r = __atomic_load_n(&t.is_rule_, __ATOMIC_RELAXED);   // every thread computes the same value
  if (r < 0) {
    r = m_.is_rule(t.name) ? 1 : 0;
    __atomic_store_n(&t.is_rule_, r, __ATOMIC_RELAXED);
  }
  return !r;
}

bool Eval::is_ground(const TP& t, const Env& env) const {
  switch (t->k) {
    case TK::Scalar: return true;
    case TK::Var: return !var_unbound(*t, env);
    case TK::Array:
    case TK::Set:
      for (auto& x : t->args)
        if (!is_ground(x, env)) return false;
      return true;
    case TK::Object:
      for (auto& kv : t->kvs)
        if (!is_ground(kv.first, env) || !is_ground(kv.second, env)) return false;
      return true;
    default: return true;   // refs / calls / comprehensions evaluate (refs may bind inner vars)
  }
}

VP Eval::eval_first(const TP& t, Env& env) {
  VP out;
  size_t mk = env.mark();
  eval_term(t, env, [&](const VP& v) {
    out = v;
    return true;
  });
  env.undo(mk);
  return out;
}

VP Eval::rule_chain(const Rule& r, Env& env) {
  VP out;
  size_t mk = env.mark();
  eval_body(r.body, 0, env, [&]() {
    out = eval_first(r.value, env);
    return (bool)out;
  });
  env.undo(mk);
  if (out) return out;
  for (auto& el : r.els) {
    eval_body(el.second, 0, env, [&]() {
      out = el.first ? eval_first(el.first, env) : v_bool(true);
      return (bool)out;
    });
    env.undo(mk);
    if (out) return out;
  }
  return nullptr;
}

VP Eval::rule_value(const std::string& name) {
  auto it = cache_has_.find(name);
  if (it != cache_has_.end()) return cache_[name];
  auto rit = m_.rules.find(name);
  if (rit == m_.rules.end()) throw RegoError{"rego_type_error: unknown rule " + name};
  const auto& rules = rit->second;
  if (++depth_ > 64) {
    --depth_;
    throw RegoError{"rego_recursion_error: rule " + name};
  }
  VP val;
  Rule::Kind kind = rules[0].kind;
  if (kind == Rule::PSet) {
    std::vector<VP> out;
    for (auto& r : rules) {
      Env env;
      eval_body(r.body, 0, env, [&]() {
        eval_term(r.key, env, [&](const VP& v) {
          out.push_back(v);
          return false;
        });
        return false;
      });
    }
    val = v_set(std::move(out));
  } else if (kind == Rule::PObj) {
    std::vector<std::pair<VP, VP>> out;
    for (auto& r : rules) {
      Env env;
      eval_body(r.body, 0, env, [&]() {
        eval_term(r.key, env, [&](const VP& k) {
          VP v = eval_first(r.value, env);
          if (v) out.emplace_back(k, v);
          return false;
        });
        return false;
      });
    }
    val = v_obj(std::move(out));
  } else if (kind == Rule::Complete) {
    VP dflt;
    for (auto& r : rules) {
      Env env;
      if (r.is_default) {
        dflt = eval_first(r.value, env);
        continue;
      }
      val = rule_chain(r, env);
      if (val) break;
    }
    if (!val) val = dflt;
  } else {
    --depth_;
    throw RegoError{"rego_type_error: " + name + " is a function"};
  }
  --depth_;
  cache_has_[name] = true;
  cache_[name] = val;
  return val;
}

const std::vector<Rule>* Eval::function_rules(const Term& t, bool* names_other_rule) const {
  // is_rule_ of a Call term: -1 unresolved, 1 a function, 2 a rule that is not a function, 0 no rule (a builtin, or unknown)
  signed char st = __atomic_load_n(&t.is_rule_, __ATOMIC_ACQUIRE);
  if (st < 0) {
    auto rit = m_.rules.find(t.name);
    const bool is_rule = rit != m_.rules.end();
    const void* fr = is_rule && rit->second[0].kind == Rule::Func ? &rit->second : nullptr;
    st = fr ? 1 : is_rule ? 2 : 0;
    __atomic_store_n(&t.rules_, fr, __ATOMIC_RELAXED);   // every thread stores the same pointer
    __atomic_store_n(&t.is_rule_, st, __ATOMIC_RELEASE);
  }
  if (names_other_rule) *names_other_rule = st == 2;
  return static_cast<const std::vector<Rule>*>(__atomic_load_n(&t.rules_, __ATOMIC_RELAXED));
}

VP Eval::call_function(const std::string& name, const std::vector<VP>& args) {
  auto rit = m_.rules.find(name);
  if (rit == m_.rules.end()) return nullptr;
  return call_function(rit->second, args);
}

VP Eval::call_function(const std::vector<Rule>& rules, const std::vector<VP>& args) {
  const std::string& name = rules[0].name;
  // pure function of scalar arguments: one evaluation per distinct argument tuple
  std::string memo_key;
  std::unordered_map<std::string, VP>* memo_map = nullptr;
  {
    auto pit = pure_of_.find(&rules);
    if (pit == pure_of_.end()) {
      auto pf = m_.pure_fn.find(name);
      pit = pure_of_.emplace(&rules, pf != m_.pure_fn.end() && pf->second).first;
    }
    bool memo = pit->second;
    for (auto& a : args) memo = memo && a && (uint8_t)a->t <= (uint8_t)VT::Str;
    if (memo) {
      const bool one_str = args.size() == 1 && args[0]->t == VT::Str;
      memo_map = &fn_memo_[reinterpret_cast<const char*>(&rules) + (one_str ? 1 : 0)];   // (two key spaces per function)
      if (one_str) {
        // by far the commonest shape (a quantity / image string): the argument itself is the key
        auto it = memo_map->find(args[0]->s);
        if (it != memo_map->end()) return it->second;
        memo_key = args[0]->s;
      } else {
        for (auto& a : args) {
          memo_key.push_back('\x01');
          memo_key.push_back((char)('0' + (int)a->t));
          if (a->t == VT::Str) memo_key += a->s;
          else if (a->t == VT::Num) memo_key += num_str(a->n);
        }
        auto it = memo_map->find(memo_key);
        if (it != memo_map->end()) return it->second;
      }
    }
  }
  if (++depth_ > 64) {
    --depth_;
    throw RegoError{"rego_recursion_error: function " + name};
  }
  VP out;
  for (auto& r : rules) {
    if (r.kind != Rule::Func || r.args.size() != args.size()) continue;
    Env env;
    // unify formals with actuals
    auto bind = [&](auto&& self, size_t i) -> bool {
      if (i == args.size()) {
        out = rule_chain(r, env);
        return (bool)out;
      }
      return unify_val(r.args[i], args[i], env, [&]() { return self(self, i + 1); });
    };
    bind(bind, 0);
    if (out) break;
  }
  --depth_;
  if (memo_map) {
    if (memo_map->size() > (1u << 16)) memo_map->clear();
    memo_map->emplace(std::move(memo_key), out);
  }
  return out;
}

bool Eval::eval_body(const std::vector<Stmt>& body, size_t i, Env& env, const EnvK& k) {
  if (i == body.size()) return k();
  const Stmt& st = body[i];
  if (st.k == Stmt::Some) return eval_body(body, i + 1, env, k);
  if (st.k == Stmt::Not) {
    bool found = false;
    size_t mk = env.mark();
    // (no copy of the statement: a TP copy would bounce the term's reference count between all flatten threads)
    eval_term(st.a, env, [&](const VP& v) {
      if (v->t == VT::False) return false;
      found = true;
      return true;
    });
    env.undo(mk);
    if (found) return false;
    return eval_body(body, i + 1, env, k);
  }
  return eval_stmt(st, env, [&]() { return eval_body(body, i + 1, env, k); });
}

bool Eval::eval_stmt(const Stmt& st, Env& env, const EnvK& k) {
  switch (st.k) {
    case Stmt::Expr: {
      size_t mk = env.mark();
      bool stop = eval_term(st.a, env, [&](const VP& v) {
        if (v->t == VT::False) return false;
        return k();
      });
      env.undo(mk);
      return stop;
    }
    case Stmt::Assign:
    case Stmt::Unify: {
      size_t mk = env.mark();
      bool stop = unify(st.a, st.b, env, k);
      env.undo(mk);
      return stop;
    }
    case Stmt::SomeIn: {
      size_t mk = env.mark();
      bool stop = eval_term(st.c, env, [&](const VP& coll) {
        auto each = [&](const VP& kk, const VP& vv) -> bool {
          size_t m2 = env.mark();
          bool s;
          if (st.a) s = unify_val(st.a, kk, env, [&]() { return unify_val(st.b, vv, env, k); });
          else s = unify_val(st.b, vv, env, k);
          env.undo(m2);
          return s;
        };
        if (coll->t == VT::Arr) {
          for (size_t j = 0; j < coll->items.size(); ++j)
            if (each(v_int((long long)j), coll->items[j])) return true;
        } else if (coll->t == VT::Set) {
          for (auto& x : coll->items)
            if (each(x, x)) return true;
        } else if (coll->t == VT::Obj) {
          for (auto& e : coll->kv)
            if (each(e.first, e.second)) return true;
        }
        return false;
      });
      env.undo(mk);
      return stop;
    }
    default: throw RegoError{"rego_unsupported: statement kind"};
  }
}

bool Eval::unify(const TP& a, const TP& b, Env& env, const EnvK& k) {
  bool ga = is_ground(a, env), gb = is_ground(b, env);
  if (ga && gb) {
    return eval_term(a, env, [&](const VP& va) {
      return eval_term(b, env, [&](const VP& vb) { return v_eq(va, vb) ? k() : false; });
    });
  }
  if (gb) return eval_term(b, env, [&](const VP& vb) { return unify_val(a, vb, env, k); });
  if (ga) return eval_term(a, env, [&](const VP& va) { return unify_val(b, va, env, k); });
  if (a->k == TK::Array && b->k == TK::Array && a->args.size() == b->args.size()) {
    std::function<bool(size_t)> rec = [&](size_t i) -> bool {
      if (i == a->args.size()) return k();
      return unify(a->args[i], b->args[i], env, [&]() { return rec(i + 1); });
    };
    return rec(0);
  }
  throw RegoError{"rego_unsafe_var_error: cannot unify two non-ground terms (line " + std::to_string(a->line) + ")"};
}

bool Eval::unify_val(const TP& pat, const VP& val, Env& env, const EnvK& k) {
  switch (pat->k) {
    case TK::Var: {
      if (const VP* b = env.find(pat->vid)) return v_eq(*b, val) ? k() : false;
      if (!var_unbound(*pat, env)) return eval_term(pat, env, [&](const VP& v) { return v_eq(v, val) ? k() : false; });
      size_t mk = env.mark();
      env.bind(pat->vid, val);
      bool s = k();
      env.undo(mk);
      return s;
    }
    case TK::Array: {
      if (val->t != VT::Arr || val->items.size() != pat->args.size()) return false;
      std::function<bool(size_t)> rec = [&](size_t i) -> bool {
        if (i == pat->args.size()) return k();
        return unify_val(pat->args[i], val->items[i], env, [&]() { return rec(i + 1); });
      };
      return rec(0);
    }
    case TK::Object: {
      if (val->t != VT::Obj || val->kv.size() != pat->kvs.size()) return false;
      std::function<bool(size_t)> rec = [&](size_t i) -> bool {
        if (i == pat->kvs.size()) return k();
        return eval_term(pat->kvs[i].first, env, [&](const VP& kv) {
          VP got = obj_get(val, kv);
          if (!got) return false;
          return unify_val(pat->kvs[i].second, got, env, [&]() { return rec(i + 1); });
        });
      };
      return rec(0);
    }
    default: return eval_term(pat, env, [&](const VP& v) { return v_eq(v, val) ? k() : false; });
  }
}

bool Eval::eval_seq(const std::vector<TP>& items, size_t i, std::vector<VP>& acc, Env& env, const EnvK& k) {
  if (i == items.size()) return k();
  return eval_term(items[i], env, [&](const VP& v) {
    acc.push_back(v);
    bool s = eval_seq(items, i + 1, acc, env, k);
    acc.pop_back();
    return s;
  });
}

bool Eval::eval_term(const TP& t, Env& env, const ValK& k) {
  switch (t->k) {
    case TK::Scalar: {
      // literals are shared by every thread that evaluates this module: hand out a thread-private copy so that
      // reference counting never bounces a cache line between flatten workers
      // (the entry pins the original node, so its address cannot be recycled while it is a key)
      static thread_local std::unordered_map<const Node*, std::pair<VP, VP>> priv;
      auto it = priv.find(t->val.get());
      if (it == priv.end()) {
        if (priv.size() > 65536) priv.clear();
        const Node& n = *t->val;
        VP c = n.t == VT::Str ? v_str(n.s) : n.t == VT::Num ? v_num(n.n) : n.t == VT::Null ? v_null() : v_bool(n.t == VT::True);
        it = priv.emplace(t->val.get(), std::make_pair(t->val, std::move(c))).first;
      }
      return k(it->second.second);
    }
    case TK::Var: {
      if (const VP* b = env.find(t->vid)) {
        VP v = *b;   // copy: the continuation may grow env and invalidate b
        return k(v);
      }
      if (t->vid == m_.vid_input) return input_ ? k(input_) : false;
      if (t->vid == m_.vid_data) return data_ ? k(data_) : false;
      if (m_.is_rule(t->name)) {
        VP v = rule_value(t->name);
        return v ? k(v) : false;
      }
      throw RegoError{"rego_unsafe_var_error: var " + t->name + " is unsafe (line " + std::to_string(t->line) + ")"};
    }
    case TK::Ref:
      return eval_term(t->head, env, [&](const VP& base) { return walk(base, t->args, 0, env, k); });
    case TK::Call: return eval_call(*t, env, k);
    case TK::Array: {
      std::vector<VP> acc;
      return eval_seq(t->args, 0, acc, env, [&]() { return k(v_arr(acc)); });
    }
    case TK::Set: {
      std::vector<VP> acc;
      return eval_seq(t->args, 0, acc, env, [&]() { return k(v_set(acc)); });
    }
    case TK::Object: {
      std::vector<std::pair<VP, VP>> acc;
      std::function<bool(size_t)> rec = [&](size_t i) -> bool {
        if (i == t->kvs.size()) return k(v_obj(acc));
        return eval_term(t->kvs[i].first, env, [&](const VP& kk) {
          return eval_term(t->kvs[i].second, env, [&](const VP& vv) {
            acc.emplace_back(kk, vv);
            bool s = rec(i + 1);
            acc.pop_back();
            return s;
          });
        });
      };
      return rec(0);
    }
    case TK::ArrCompr:
    case TK::SetCompr: {
      std::vector<VP> out;
      size_t mk = env.mark();
      eval_body(t->body, 0, env, [&]() {
        eval_term(t->value, env, [&](const VP& v) {
          out.push_back(v);
          return false;
        });
        return false;
      });
      env.undo(mk);
      return k(t->k == TK::ArrCompr ? v_arr(std::move(out)) : v_set(std::move(out)));
    }
    case TK::ObjCompr: {
      std::vector<std::pair<VP, VP>> out;
      size_t mk = env.mark();
      eval_body(t->body, 0, env, [&]() {
        eval_term(t->key, env, [&](const VP& kk) {
          VP vv = eval_first(t->value, env);
          if (vv) out.emplace_back(kk, vv);
          return false;
        });
        return false;
      });
      env.undo(mk);
      return k(v_obj(std::move(out)));
    }
  }
  return false;
}

bool Eval::walk(const VP& cur, const std::vector<TP>& path, size_t i, Env& env, const ValK& k) {
  if (i == path.size()) return k(cur);
  const TP& p = path[i];
  if (var_unbound(*p, env)) {
    auto step = [&](const VP& kk, const VP& vv) -> bool {
      size_t mk = env.mark();
      env.bind(p->vid, kk);
      bool s = walk(vv, path, i + 1, env, k);
      env.undo(mk);
      return s;
    };
    if (cur->t == VT::Arr) {
      for (size_t j = 0; j < cur->items.size(); ++j)
        if (step(v_int((long long)j), cur->items[j])) return true;
    } else if (cur->t == VT::Obj) {
      for (auto& e : cur->kv)
        if (step(e.first, e.second)) return true;
    } else if (cur->t == VT::Set) {
      for (auto& x : cur->items)
        if (step(x, x)) return true;
    }
    return false;
  }
  if (!is_ground(p, env)) {
    // pattern key, e.g. general_violation[{"msg": msg, "field": "containers"}]
    if (cur->t == VT::Set) {
      for (auto& x : cur->items) {
        size_t mk = env.mark();
        bool s = unify_val(p, x, env, [&]() { return walk(x, path, i + 1, env, k); });
        env.undo(mk);
        if (s) return true;
      }
    } else if (cur->t == VT::Obj) {
      for (auto& e : cur->kv) {
        size_t mk = env.mark();
        bool s = unify_val(p, e.first, env, [&]() { return walk(e.second, path, i + 1, env, k); });
        env.undo(mk);
        if (s) return true;
      }
    }
    return false;
  }
  return eval_term(p, env, [&](const VP& kv) {
    if (cur->t == VT::Obj) {
      VP got = obj_get(cur, kv);
      return got ? walk(got, path, i + 1, env, k) : false;
    }
    if (cur->t == VT::Arr) {
      int64_t ix;
      if (kv->t == VT::Num && num_fits_i64(kv->n, &ix) && ix >= 0 && (size_t)ix < cur->items.size())
        return walk(cur->items[ix], path, i + 1, env, k);
      return false;
    }
    if (cur->t == VT::Set) {
      VP got = set_find(cur, kv);
      return got ? walk(got, path, i + 1, env, k) : false;
    }
    return false;
  });
}

bool Eval::eval_call(const Term& t, Env& env, const ValK& k) {
  const auto* frules = function_rules(t);
  bool user = frules != nullptr;
  size_t nargs = t.args.size();
  const TP* out_pat = nullptr;
  if (user && nargs == (*frules)[0].args.size() + 1) {
    out_pat = &t.args.back();
    --nargs;
  }
  std::vector<VP> acc;
  acc.reserve(nargs);
  auto rec = [&](auto&& self, size_t i) -> bool {
    if (i == nargs) {
      VP v;
      if (user) v = call_function(*frules, acc);
      else {
        bool known = true;
        v = call_builtin(t.name, acc, &known);
        if (!known) throw RegoError{"rego_type_error: undefined function " + t.name + " (line " + std::to_string(t.line) + ")"};
      }
      if (!v) return false;
      if (out_pat) return unify_val(*out_pat, v, env, [&]() { return k(v_bool(true)); });
      return k(v);
    }
    return eval_term(t.args[i], env, [&](const VP& v) {
      acc.push_back(v);
      bool s = self(self, i + 1);
      acc.pop_back();
      return s;
    });
  };
  return rec(rec, 0);
}

}  // namespace gk
