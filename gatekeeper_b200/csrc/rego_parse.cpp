// Rego subset parser (v0 and v1 rule syntax).  See rego.hpp for scope.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <set>

#include "rego.hpp"

namespace gk {

namespace {

enum class TT : uint8_t { Num, Str, Id, Op, Eof };
struct Tok {
  TT k;
  std::string s;
  VP v;
  int line;
  size_t pos, end;
};

[[noreturn]] void perr(const std::string& m, int line) {
  throw RegoError{"rego_parse_error: " + m + " (line " + std::to_string(line) + ")"};
}

std::vector<Tok> lex(const std::string& src) {
  std::vector<Tok> out;
  size_t i = 0, n = src.size();
  int line = 1;
  auto prev_is_operand = [&]() {
    if (out.empty() || out.back().line != line) return false;
    const Tok& t = out.back();
    if (t.k == TT::Num || t.k == TT::Str || t.k == TT::Id) return true;
    return t.k == TT::Op && (t.s == ")" || t.s == "]" || t.s == "}");
  };
  while (i < n) {
    char c = src[i];
    if (c == '\n') {
      ++line;
      ++i;
      continue;
    }
    if (c == ' ' || c == '\t' || c == '\r') {
      ++i;
      continue;
    }
    if (c == '#') {
      while (i < n && src[i] != '\n') ++i;
      continue;
    }
    size_t st = i;
    if ((c >= '0' && c <= '9') || (c == '-' && i + 1 < n && src[i + 1] >= '0' && src[i + 1] <= '9' && !prev_is_operand())) {
      if (c == '-') ++i;
      bool isint = true;
      while (i < n && src[i] >= '0' && src[i] <= '9') ++i;
      if (i + 1 < n && src[i] == '.' && src[i + 1] >= '0' && src[i + 1] <= '9') {
        isint = false;
        ++i;
        while (i < n && src[i] >= '0' && src[i] <= '9') ++i;
      }
      if (i < n && (src[i] == 'e' || src[i] == 'E')) {
        size_t j = i + 1;
        if (j < n && (src[j] == '+' || src[j] == '-')) ++j;
        if (j < n && src[j] >= '0' && src[j] <= '9') {
          isint = false;
          i = j;
          while (i < n && src[i] >= '0' && src[i] <= '9') ++i;
        }
      }
      std::string txt = src.substr(st, i - st);
      VP v;
      try {
        v = json_parse(txt.data(), txt.size());
      } catch (JsonError&) {
        perr("bad number " + txt, line);
      }
      (void)isint;
      out.push_back({TT::Num, txt, v, line, st, i});
      continue;
    }
    if (c == '"') {
      size_t j = i + 1;
      while (j < n && src[j] != '"') {
        if (src[j] == '\\') ++j;
        if (j < n && src[j] == '\n') perr("newline in string", line);
        ++j;
      }
      if (j >= n) perr("unterminated string", line);
      std::string txt = src.substr(i, j + 1 - i);
      VP v;
      try {
        v = json_parse(txt.data(), txt.size());
      } catch (JsonError&) {
        perr("bad string literal", line);
      }
      out.push_back({TT::Str, v->s, v, line, st, j + 1});
      i = j + 1;
      continue;
    }
    if (c == '`') {
      size_t j = src.find('`', i + 1);
      if (j == std::string::npos) perr("unterminated raw string", line);
      std::string body = src.substr(i + 1, j - i - 1);
      out.push_back({TT::Str, body, v_str(body), line, st, j + 1});
      for (char ch : body)
        if (ch == '\n') ++line;
      i = j + 1;
      continue;
    }
    if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_') {
      while (i < n && ((src[i] >= 'a' && src[i] <= 'z') || (src[i] >= 'A' && src[i] <= 'Z') || (src[i] >= '0' && src[i] <= '9') || src[i] == '_')) ++i;
      out.push_back({TT::Id, src.substr(st, i - st), nullptr, line, st, i});
      continue;
    }
    static const char* two[] = {":=", "==", "!=", "<=", ">="};
    bool got = false;
    for (const char* t : two) {
      if (i + 1 < n && src[i] == t[0] && src[i + 1] == t[1]) {
        out.push_back({TT::Op, t, nullptr, line, st, i + 2});
        i += 2;
        got = true;
        break;
      }
    }
    if (got) continue;
    if (strchr("=<>+-*/%&|[]{}().,;:", c)) {
      out.push_back({TT::Op, std::string(1, c), nullptr, line, st, i + 1});
      ++i;
      continue;
    }
    perr(std::string("illegal character '") + c + "'", line);
  }
  out.push_back({TT::Eof, "", nullptr, line, n, n});
  return out;
}

int prec_of(const std::string& op) {
  if (op == "==" || op == "!=" || op == "<" || op == "<=" || op == ">" || op == ">=" || op == "in") return 1;
  if (op == "|") return 2;
  if (op == "&") return 3;
  if (op == "+" || op == "-") return 4;
  if (op == "*" || op == "/" || op == "%") return 5;
  return 0;
}

const char* infix_name(const std::string& op) {
  if (op == "==") return "equal";
  if (op == "!=") return "neq";
  if (op == "<") return "lt";
  if (op == "<=") return "lte";
  if (op == ">") return "gt";
  if (op == ">=") return "gte";
  if (op == "+") return "plus";
  if (op == "-") return "minus";
  if (op == "*") return "mul";
  if (op == "/") return "div";
  if (op == "%") return "rem";
  if (op == "&") return "and";
  if (op == "|") return "or";
  return "";
}

bool is_keyword(const std::string& s) {
  static const char* kw[] = {"not", "some", "default", "package", "import", "else", "with", "as"};
  for (auto k : kw)
    if (s == k) return true;
  return false;
}

struct Parser {
  std::vector<Tok> t;
  size_t i = 0;
  Module& m;
  int wild = 0;
  // template libs (`spec.targets[].libs`): every lib is parsed into the SAME module, its rules renamed to
  // "data.<package>.<rule>"; an `import data.lib.x [as y]` makes `y.rule` / `y.fn(...)` mean "data.lib.x.rule".
  std::string prefix;                              // "data.lib.helpers." while a lib is parsed, "" for the entry point
  const std::set<std::string>* own_rules = nullptr;   // the lib's own rule names (unqualified)
  const std::set<std::string>* lib_pkgs = nullptr;    // "data.lib.helpers", ... : every lib package of the template
  std::map<std::string, std::string> aliases;      // import alias -> "data.lib.helpers."
  bool lenient_imports = false;                    // pre-scan: take any data.lib import
  explicit Parser(const std::string& src, Module& mod) : t(lex(src)), m(mod) {}

  std::string resolve(const std::string& n) const {
    if (!prefix.empty() && own_rules && own_rules->count(n)) return prefix + n;
    size_t dot = n.find('.');
    if (dot != std::string::npos) {
      auto it = aliases.find(n.substr(0, dot));
      if (it != aliases.end()) return it->second + n.substr(dot + 1);
    }
    return n;
  }

  const Tok& peek(size_t k = 0) const { return t[std::min(i + k, t.size() - 1)]; }
  const Tok& next() { return t[i < t.size() - 1 ? i++ : i]; }
  bool at(const char* v) const {
    const Tok& x = peek();
    return (x.k == TT::Op || x.k == TT::Id) && x.s == v;
  }
  bool accept(const char* v) {
    if (at(v)) {
      ++i;
      return true;
    }
    return false;
  }
  void expect(const char* v) {
    if (!accept(v)) perr(std::string("expected '") + v + "' got '" + peek().s + "'", peek().line);
  }
  bool same_line() const { return i > 0 && peek().line == t[i - 1].line; }
  bool adjacent() const { return i > 0 && peek().pos == t[i - 1].end; }

  std::shared_ptr<Term> mk(TK k, int line) {
    auto x = std::make_shared<Term>();
    x->k = k;
    x->line = line;
    return x;
  }
  TP scalar(VP v, int line) {
    auto x = mk(TK::Scalar, line);
    x->val = std::move(v);
    return x;
  }
  TP var(const std::string& n0, int line) {
    const std::string n = resolve(n0);
    auto x = mk(TK::Var, line);
    x->name = n;
    x->vid = m.intern(n);
    return x;
  }
  TP call(const std::string& n0, std::vector<TP> args, int line) {
    const std::string n = resolve(n0);
    auto x = mk(TK::Call, line);
    x->name = n;
    x->args = std::move(args);
    return x;
  }
  TP ref_append(const TP& base, TP idx) {
    if (idx->k == TK::Scalar && idx->val->t == VT::Str) {
      // <alias>.<rule> and data.lib.<pkg>.<rule> name a lib rule
      if (base->k == TK::Var) {
        auto it = aliases.find(base->name);
        if (it != aliases.end()) return var(it->second + idx->val->s, base->line);
      }
      if (lib_pkgs && base->k == TK::Ref && base->head->k == TK::Var && base->head->vid == m.vid_data) {
        std::string pkg = "data";
        bool plain = true;
        for (auto& a : base->args) {
          if (a->k != TK::Scalar || a->val->t != VT::Str) {
            plain = false;
            break;
          }
          pkg += "." + a->val->s;
        }
        if (plain && lib_pkgs->count(pkg)) return var(pkg + "." + idx->val->s, base->line);
      }
    }
    if (base->k == TK::Ref) {
      auto x = std::make_shared<Term>(*base);
      x->args.push_back(std::move(idx));
      return x;
    }
    auto x = mk(TK::Ref, base->line);
    x->head = base;
    x->args.push_back(std::move(idx));
    return x;
  }
  static std::string dotted(const TP& b, int line) {
    if (b->k == TK::Var) return b->name;
    std::string s = b->head->name;
    for (auto& a : b->args) {
      if (a->k != TK::Scalar || a->val->t != VT::Str) perr("bad function name", line);
      s += "." + a->val->s;
    }
    return s;
  }

  void parse_module() {
    bool have_pkg = false;
    while (peek().k != TT::Eof) {
      if (at("package")) {
        next();
        std::string p = next().s;
        while (accept(".")) p += "." + next().s;
        m.package = p;
        have_pkg = true;
      } else if (at("import")) {
        int line = next().line;
        std::string p = next().s;
        std::string first = p;
        while (same_line() && accept(".")) p += "." + next().s;
        std::string alias = p.substr(p.rfind('.') + 1);
        if (same_line() && accept("as")) alias = next().s;
        if (first == "future" || first == "rego") {
        } else if (p.rfind("data.lib.", 0) == 0 && (lenient_imports || (lib_pkgs && lib_pkgs->count(p)))) {
          aliases[alias] = p + ".";
        } else if (p.rfind("data.lib.", 0) == 0 || p == "data.lib") {
          throw RegoError{"rego_compile_error: import " + p + ": the template has no lib with that package (line " + std::to_string(line) + ")"};
        } else {
          throw RegoError{"rego_unsupported: import " + p + " (line " + std::to_string(line) + ")"};
        }
      } else {
        Rule r = parse_rule();
        m.rules[r.name].push_back(std::move(r));
      }
    }
    if (!have_pkg) throw RegoError{"rego_parse_error: package expected"};
  }

  Rule parse_rule() {
    Rule r;
    r.is_default = accept("default");
    const Tok& nt = next();
    if (nt.k != TT::Id || is_keyword(nt.s)) perr("unexpected '" + nt.s + "'", nt.line);
    r.name = prefix + nt.s;
    r.line = nt.line;
    if (same_line() && at("(") && adjacent()) {
      next();
      while (!at(")")) {
        r.args.push_back(parse_term());
        if (!accept(",")) break;
      }
      expect(")");
      r.kind = Rule::Func;
    } else if (same_line() && at("[") && adjacent()) {
      next();
      r.key = parse_term();
      expect("]");
      r.kind = Rule::PSet;
    }
    if (at("contains")) {
      next();
      r.key = parse_term();
      r.kind = Rule::PSet;
    }
    if (at("=") || at(":=")) {
      next();
      r.value = parse_term();
      if (r.kind == Rule::PSet) r.kind = Rule::PObj;
    }
    bool has_if = accept("if");
    if (at("{")) {
      next();
      r.body = parse_body("}");
      expect("}");
      r.has_body = true;
    } else if (has_if) {
      r.body.push_back(parse_stmt());
      r.has_body = true;
    }
    while (at("else")) {
      next();
      TP ev;
      if (at("=") || at(":=")) {
        next();
        ev = parse_term();
      }
      accept("if");
      std::vector<Stmt> eb;
      if (at("{")) {
        next();
        eb = parse_body("}");
        expect("}");
      }
      r.els.emplace_back(ev, std::move(eb));
    }
    if (!r.value && (r.kind == Rule::Complete || r.kind == Rule::Func)) r.value = scalar(v_bool(true), r.line);
    return r;
  }

  std::vector<Stmt> parse_body(const char* closer) {
    std::vector<Stmt> out;
    while (!at(closer)) {
      if (peek().k == TT::Eof) perr("unexpected eof in body", peek().line);
      out.push_back(parse_stmt());
      if (accept(";")) continue;
      if (!at(closer) && same_line()) perr("unexpected '" + peek().s + "'", peek().line);
    }
    return out;
  }

  Stmt parse_stmt() {
    Stmt s;
    s.line = peek().line;
    if (at("some")) {
      next();
      std::vector<TP> names;
      names.push_back(parse_term(0, true));
      while (accept(",")) names.push_back(parse_term(0, true));
      if (accept("in")) {
        s.k = Stmt::SomeIn;
        s.c = parse_term();
        if (names.size() == 1) s.b = names[0];
        else {
          s.a = names[0];
          s.b = names[1];
        }
        return s;
      }
      s.k = Stmt::Some;
      {   // the declared names are kept (as an array term) for the scoping pass below; evaluation ignores the statement
        auto arr = mk(TK::Array, s.line);
        arr->args = names;
        s.a = arr;
      }
      return s;
    }
    if (at("every") && peek(1).k == TT::Id && !is_keyword(peek(1).s)) {
      next();
      TP first = parse_term(0, true), second;
      if (accept(",")) second = parse_term(0, true);
      expect("in");
      s.k = Stmt::Every;
      s.c = parse_term();
      if (second) s.a = first, s.b = second;
      else s.b = first;
      expect("{");
      s.body = parse_body("}");
      expect("}");
      return s;
    }
    if (at("not")) {
      next();
      Stmt inner = parse_expr();
      if (inner.k != Stmt::Expr) {
        // `not a = b` : negated unification, treat as not (a == b) when ground
        s.k = Stmt::Not;
        s.a = call("equal", {inner.a, inner.b}, s.line);
        return s;
      }
      s.k = Stmt::Not;
      s.a = inner.a;
      return s;
    }
    s = parse_expr();
    if (same_line() && at("with")) throw RegoError{"rego_unsupported: `with` modifier (line " + std::to_string(s.line) + ")"};
    return s;
  }

  Stmt parse_expr() {
    Stmt s;
    s.line = peek().line;
    TP lhs = parse_term();
    if (same_line() && (at(":=") || at("="))) {
      bool assign = next().s == ":=";
      s.k = assign ? Stmt::Assign : Stmt::Unify;
      s.a = lhs;
      s.b = parse_term();
      return s;
    }
    s.k = Stmt::Expr;
    s.a = lhs;
    return s;
  }

  TP parse_term(int prec = 0, bool no_in = false, bool no_bar = false) {
    TP lhs = parse_unary();
    while (same_line()) {
      const Tok& x = peek();
      std::string op;
      if (x.k == TT::Op && prec_of(x.s) && !(no_bar && x.s == "|")) op = x.s;
      else if (x.k == TT::Id && x.s == "in" && !no_in) op = "in";
      else break;
      int p = prec_of(op);
      if (p <= prec) break;
      int line = next().line;
      TP rhs = parse_term(p, no_in, no_bar);
      if (op == "in") lhs = call("internal.member_2", {lhs, rhs}, line);
      else lhs = call(infix_name(op), {lhs, rhs}, line);
    }
    return lhs;
  }

  TP parse_unary() {
    if (peek().k == TT::Op && peek().s == "-") {
      int line = next().line;
      TP x = parse_unary();
      return call("minus", {scalar(v_int(0), line), x}, line);
    }
    return parse_postfix(parse_primary());
  }

  TP parse_postfix(TP base) {
    while (same_line()) {
      if (at(".") ) {
        next();
        const Tok& f = next();
        if (f.k != TT::Id) perr("bad ref", f.line);
        base = ref_append(base, scalar(v_str(f.s), f.line));
      } else if (at("[") && adjacent()) {
        next();
        TP idx = parse_term();
        expect("]");
        base = ref_append(base, idx);
      } else if (at("(") && adjacent() && (base->k == TK::Var || base->k == TK::Ref)) {
        int line = next().line;
        std::vector<TP> args;
        while (!at(")")) {
          args.push_back(parse_term());
          if (!accept(",")) break;
        }
        expect(")");
        base = call(dotted(base, line), std::move(args), line);
      } else {
        break;
      }
    }
    return base;
  }

  TP parse_primary() {
    const Tok& x = next();
    int line = x.line;
    if (x.k == TT::Num || x.k == TT::Str) return scalar(x.v, line);
    if (x.k == TT::Id) {
      if (x.s == "true") return scalar(v_bool(true), line);
      if (x.s == "false") return scalar(v_bool(false), line);
      if (x.s == "null") return scalar(v_null(), line);
      if (x.s == "_") return var("$w" + std::to_string(++wild), line);
      if (x.s == "set" && at("(") && peek(1).s == ")") {
        next();
        next();
        return mk(TK::Set, line);
      }
      if (is_keyword(x.s)) perr("unexpected keyword '" + x.s + "'", line);
      return var(x.s, line);
    }
    if (x.k == TT::Op) {
      if (x.s == "(") {
        TP e = parse_term();
        expect(")");
        return e;
      }
      if (x.s == "[") {
        if (accept("]")) return mk(TK::Array, line);
        TP first = parse_term(0, false, true);
        if (accept("|")) {
          auto c = mk(TK::ArrCompr, line);
          c->value = first;
          c->body = parse_body("]");
          expect("]");
          return c;
        }
        auto a = mk(TK::Array, line);
        a->args.push_back(first);
        while (accept(",")) {
          if (at("]")) break;
          a->args.push_back(parse_term());
        }
        expect("]");
        return a;
      }
      if (x.s == "{") {
        if (accept("}")) return mk(TK::Object, line);
        TP first = parse_term(0, false, true);
        if (accept(":")) {
          TP val = parse_term(0, false, true);
          if (accept("|")) {
            auto c = mk(TK::ObjCompr, line);
            c->key = first;
            c->value = val;
            c->body = parse_body("}");
            expect("}");
            return c;
          }
          auto o = mk(TK::Object, line);
          o->kvs.emplace_back(first, val);
          while (accept(",")) {
            if (at("}")) break;
            TP k = parse_term();
            expect(":");
            o->kvs.emplace_back(k, parse_term());
          }
          expect("}");
          return o;
        }
        if (accept("|")) {
          auto c = mk(TK::SetCompr, line);
          c->value = first;
          c->body = parse_body("}");
          expect("}");
          return c;
        }
        auto s = mk(TK::Set, line);
        s->args.push_back(first);
        while (accept(",")) {
          if (at("}")) break;
          s->args.push_back(parse_term());
        }
        expect("}");
        return s;
      }
    }
    perr("unexpected token '" + x.s + "'", line);
  }
};

void collect_vars(const TP& t, std::vector<int>& out);
void collect_vars_body(const std::vector<Stmt>& b, std::vector<int>& out) {
  for (auto& s : b) {
    if (s.a) collect_vars(s.a, out);
    if (s.b) collect_vars(s.b, out);
    if (s.c) collect_vars(s.c, out);
  }
}
void collect_vars(const TP& t, std::vector<int>& out) {
  if (!t) return;
  if (t->k == TK::Var) out.push_back(t->vid);
  if (t->head) collect_vars(t->head, out);
  for (auto& a : t->args) collect_vars(a, out);
  for (auto& kv : t->kvs) {
    collect_vars(kv.first, out);
    collect_vars(kv.second, out);
  }
  collect_vars(t->key, out);
  collect_vars(t->value, out);
  collect_vars_body(t->body, out);
}

// ---- `every k, v in coll { body }` (OPA v1 keyword: the body holds for every element; an empty domain is true, an undefined
// one undefined) is rewritten into constructs both evaluators already have:
//       $evN := coll
//       count([1 | some $ekN, $exN in $evN; not $everyN($ekN, $exN, <captured>)]) == 0
//       $everyN(k, v, <captured>) { body }
// where <captured> are the variables of the body that the enclosing bodies have bound before the statement.
struct EveryDesugar {
  Module& m;
  int counter = 0;
  std::vector<std::pair<std::string, Rule>> new_rules;

  static void var_terms(const TP& t, std::vector<TP>& out) {
    if (!t) return;
    if (t->k == TK::Var) out.push_back(t);
    var_terms(t->head, out);
    for (auto& a : t->args) var_terms(a, out);
    for (auto& kv : t->kvs) var_terms(kv.first, out), var_terms(kv.second, out);
    var_terms(t->key, out);
    var_terms(t->value, out);
    for (auto& st : t->body) var_terms(st, out);
  }
  static void var_terms(const Stmt& st, std::vector<TP>& out) {
    var_terms(st.a, out), var_terms(st.b, out), var_terms(st.c, out);
    for (auto& b : st.body) var_terms(b, out);
  }
  std::shared_ptr<Term> mk(TK k, int line) {
    auto x = std::make_shared<Term>();
    x->k = k;
    x->line = line;
    return x;
  }
  TP var(const std::string& n, int line) {
    auto x = mk(TK::Var, line);
    x->name = n;
    x->vid = m.intern(n);
    return x;
  }
  TP call(const std::string& n, std::vector<TP> args, int line) {
    auto x = mk(TK::Call, line);
    x->name = n;
    x->args = std::move(args);
    return x;
  }
  TP num(long long v, int line) {
    auto x = mk(TK::Scalar, line);
    x->val = v_int(v);
    return x;
  }

  TP fix_term(const TP& t, const std::vector<int>& bound) {
    if (!t || t->k == TK::Scalar || t->k == TK::Var) return t;
    auto x = std::make_shared<Term>(*t);
    x->head = fix_term(t->head, bound);
    for (auto& a : x->args) a = fix_term(a, bound);
    for (auto& kv : x->kvs) kv = {fix_term(kv.first, bound), fix_term(kv.second, bound)};
    x->key = fix_term(t->key, bound);
    x->value = fix_term(t->value, bound);
    if (!t->body.empty()) x->body = fix_body(t->body, bound);
    return x;
  }
  std::vector<Stmt> fix_body(const std::vector<Stmt>& body, std::vector<int> bound) {
    std::vector<Stmt> out;
    for (auto& st0 : body) {
      Stmt st = st0;
      st.a = fix_term(st.a, bound), st.b = fix_term(st.b, bound), st.c = fix_term(st.c, bound);
      if (st.k == Stmt::Every) {
        const int id = ++counter, line = st.line;
        for (const TP* p : {&st.a, &st.b})
          if (*p && (*p)->k != TK::Var) throw RegoError{"rego_unsupported: `every` with a non-variable key / value pattern (line " + std::to_string(line) + ")"};
        std::vector<int> inner = bound;
        if (st.a) inner.push_back(st.a->vid);
        inner.push_back(st.b->vid);
        std::vector<Stmt> fbody = fix_body(st.body, inner);
        // captured: variables of the body bound by the enclosing bodies (not the key / value, not documents, not rules)
        std::vector<TP> used, caps;
        for (auto& b : fbody) var_terms(b, used);
        for (auto& u : used) {
          if (u->vid == m.vid_input || u->vid == m.vid_data || m.is_rule(u->name)) continue;
          if ((st.a && u->vid == st.a->vid) || u->vid == st.b->vid) continue;
          if (std::find(bound.begin(), bound.end(), u->vid) == bound.end()) continue;
          bool dup = false;
          for (auto& c : caps) dup = dup || c->vid == u->vid;
          if (!dup) caps.push_back(u);
        }
        const std::string sfx = std::to_string(id), fn = "$every" + sfx;
        Rule r;
        r.kind = Rule::Func;
        r.name = fn;
        r.line = line;
        r.args.push_back(st.a ? st.a : var("$eku" + sfx, line));
        r.args.push_back(st.b);
        for (auto& c : caps) r.args.push_back(c);
        r.body = fbody;
        r.has_body = true;
        auto tv = mk(TK::Scalar, line);
        tv->val = v_bool(true);
        r.value = tv;
        new_rules.emplace_back(fn, std::move(r));
        // $evN := coll
        Stmt s1;
        s1.k = Stmt::Assign;
        s1.line = line;
        s1.a = var("$ev" + sfx, line);
        s1.b = st.c;
        // count([1 | some $ekN, $exN in $evN; not $everyN($ekN, $exN, caps...)]) == 0
        TP ek = var("$ek" + sfx, line), ex = var("$ex" + sfx, line);
        Stmt it;
        it.k = Stmt::SomeIn;
        it.line = line;
        it.a = ek, it.b = ex, it.c = s1.a;
        std::vector<TP> cargs{ek, ex};
        for (auto& c : caps) cargs.push_back(c);
        Stmt neg;
        neg.k = Stmt::Not;
        neg.line = line;
        neg.a = call(fn, cargs, line);
        auto compr = mk(TK::ArrCompr, line);
        compr->value = num(1, line);
        compr->body = {it, neg};
        Stmt s2;
        s2.k = Stmt::Expr;
        s2.line = line;
        s2.a = call("equal", {call("count", {TP(compr)}, line), num(0, line)}, line);
        out.push_back(s1);
        out.push_back(s2);
        bound.push_back(s1.a->vid);
        continue;
      }
      out.push_back(st);
      std::vector<TP> vs;
      var_terms(st, vs);
      for (auto& v : vs) bound.push_back(v->vid);
    }
    return out;
  }
  void run() {
    for (auto& kv : m.rules)
      for (auto& r : kv.second) {
        std::vector<int> bound;
        std::vector<TP> vs;
        for (auto& a : r.args) var_terms(a, vs);
        for (auto& v : vs) bound.push_back(v->vid);
        r.body = fix_body(r.body, bound);
        std::vector<TP> bv;
        for (auto& st : r.body) var_terms(st, bv);
        std::vector<int> after = bound;
        for (auto& v : bv) after.push_back(v->vid);
        r.key = fix_term(r.key, after), r.value = fix_term(r.value, after);
        for (auto& e : r.els) {
          e.second = fix_body(e.second, bound);
          e.first = fix_term(e.first, after);
        }
      }
    for (auto& nr : new_rules) m.rules[nr.first].push_back(std::move(nr.second));
  }
};

// ---- scoping of comprehension locals.  A variable declared with `:=` / `some` inside a comprehension is local to it and
// SHADOWS a variable of the same name in the enclosing body (`c := containers[i]; all([ok(c) | c := containers[_]])` tests every
// container, not just the outer one).  The evaluators bind by name, so every such local gets a name of its own.
struct ComprScoper {
  Module& m;
  int counter = 0;
  using Ren = std::map<int, std::pair<int, std::string>>;

  static void pattern_vars(const TP& t, std::vector<const Term*>& out) {
    if (!t) return;
    if (t->k == TK::Var) out.push_back(t.get());
    else if (t->k == TK::Array || t->k == TK::Set)
      for (auto& a : t->args) pattern_vars(a, out);
    else if (t->k == TK::Object)
      for (auto& kv : t->kvs) pattern_vars(kv.second, out);
  }
  TP rename(const TP& t, const Ren& ren) {
    if (!t) return t;
    auto x = std::make_shared<Term>(*t);
    x->is_rule_ = -1;
    x->rules_ = nullptr;
    if (t->k == TK::Var) {
      auto it = ren.find(t->vid);
      if (it == ren.end()) return t;
      x->vid = it->second.first;
      x->name = it->second.second;
      return x;
    }
    x->head = rename(t->head, ren);
    for (auto& a : x->args) a = rename(a, ren);
    for (auto& kv : x->kvs) kv = {rename(kv.first, ren), rename(kv.second, ren)};
    x->key = rename(t->key, ren);
    x->value = rename(t->value, ren);
    for (auto& st : x->body) st = rename(st, ren);
    return x;
  }
  Stmt rename(const Stmt& s, const Ren& ren) {
    Stmt o = s;
    o.a = rename(s.a, ren), o.b = rename(s.b, ren), o.c = rename(s.c, ren);
    return o;
  }
  TP fix(const TP& t) {
    if (!t) return t;
    if (t->k == TK::Scalar || t->k == TK::Var) return t;
    auto x = std::make_shared<Term>(*t);
    x->head = fix(t->head);
    for (auto& a : x->args) a = fix(a);
    for (auto& kv : x->kvs) kv = {fix(kv.first), fix(kv.second)};
    x->key = fix(t->key);
    x->value = fix(t->value);
    for (auto& st : x->body) st = fix(st);
    if (t->k != TK::ArrCompr && t->k != TK::SetCompr && t->k != TK::ObjCompr) return x;
    std::vector<const Term*> decl;
    for (auto& st : x->body) {
      if (st.k == Stmt::Assign || st.k == Stmt::Some) pattern_vars(st.a, decl);
      if (st.k == Stmt::SomeIn) pattern_vars(st.a, decl), pattern_vars(st.b, decl);
    }
    Ren ren;
    for (const Term* v : decl) {
      if (v->name.empty() || v->name[0] == '$' || ren.count(v->vid)) continue;
      std::string nn = v->name + "$" + std::to_string(++counter);
      ren[v->vid] = {m.intern(nn), nn};
    }
    if (ren.empty()) return x;
    return rename(TP(x), ren);
  }
  Stmt fix(const Stmt& s) {
    Stmt o = s;
    o.a = fix(s.a), o.b = fix(s.b), o.c = fix(s.c);
    return o;
  }
  void run() {
    for (auto& kv : m.rules)
      for (auto& r : kv.second) {
        for (auto& a : r.args) a = fix(a);
        r.key = fix(r.key), r.value = fix(r.value);
        for (auto& st : r.body) st = fix(st);
        for (auto& e : r.els) {
          e.first = fix(e.first);
          for (auto& st : e.second) st = fix(st);
        }
      }
  }
};

// ---- body ordering.  OPA's compiler reorders the expressions of a body so that every variable is bound before it is needed
// (ast/compile.go reorderBodyForSafety); templates rely on it, e.g. pkg/gator/fixtures/fixtures.go:461
//   selectors := [s | s = concat(":", [key, val]); val = obj.spec.selector[key]]
// The evaluator and the lowering walk a body left to right, so bodies are put into a safe order once, at load.  A body that is
// already safe in its written order is left alone.
struct BodyOrder {
  Module& m;
  using Set = std::set<int>;
  std::map<int, std::string> names;

  bool plain(const TP& t) {
    if (!t || t->k != TK::Var || t->vid == m.vid_input || t->vid == m.vid_data || m.is_rule(t->name)) return false;
    names.emplace(t->vid, t->name);
    return true;
  }
  // `deep` = false: the variables a body itself names (the locals of its comprehensions are not visible outside them)
  void all_vars(const TP& t, Set& out, bool deep = true) {
    if (!t) return;
    if (plain(t)) out.insert(t->vid);
    if (!deep && (t->k == TK::ArrCompr || t->k == TK::SetCompr || t->k == TK::ObjCompr)) return;
    all_vars(t->head, out, deep);
    for (auto& a : t->args) all_vars(a, out, deep);
    for (auto& kv : t->kvs) all_vars(kv.first, out, deep), all_vars(kv.second, out, deep);
    all_vars(t->key, out, deep);
    all_vars(t->value, out, deep);
    for (auto& st : t->body) all_vars(st, out, deep);
  }
  void all_vars(const Stmt& st, Set& out, bool deep = true) { all_vars(st.a, out, deep), all_vars(st.b, out, deep), all_vars(st.c, out, deep); }

  // variables needed to evaluate `t` as a value / variables that doing so binds (reference index positions)
  void value_use(const TP& t, const Set& scope, Set& need, Set& out) {
    if (!t) return;
    switch (t->k) {
      case TK::Scalar: return;
      case TK::Var:
        if (plain(t)) need.insert(t->vid);
        return;
      case TK::Ref:
        if (plain(t->head)) need.insert(t->head->vid);
        else value_use(t->head, scope, need, out);
        for (auto& a : t->args) {
          if (plain(a)) out.insert(a->vid);
          else value_use(a, scope, need, out);
        }
        return;
      case TK::ArrCompr:
      case TK::SetCompr:
      case TK::ObjCompr: {
        Set inner;
        all_vars(t->key, inner);
        all_vars(t->value, inner);
        for (auto& st : t->body) all_vars(st, inner);
        for (int v : inner)
          if (scope.count(v)) need.insert(v);   // its closure: variables of the enclosing bodies
        return;
      }
      default:
        for (auto& a : t->args) value_use(a, scope, need, out);
        for (auto& kv : t->kvs) value_use(kv.first, scope, need, out), value_use(kv.second, scope, need, out);
        return;
    }
  }
  void pattern_use(const TP& t, const Set& scope, Set& need, Set& out) {
    if (!t) return;
    if (plain(t)) {
      out.insert(t->vid);
    } else if (t->k == TK::Array) {
      for (auto& a : t->args) pattern_use(a, scope, need, out);
    } else if (t->k == TK::Object) {
      for (auto& kv : t->kvs) value_use(kv.first, scope, need, out), pattern_use(kv.second, scope, need, out);
    } else {
      value_use(t, scope, need, out);
    }
  }
  struct Opt {
    Set need, out;
  };
  // the statement can run once ONE option's `need` is bound
  std::vector<Opt> options(const Stmt& st, const Set& scope) {
    std::vector<Opt> o;
    switch (st.k) {
      case Stmt::Some: o.emplace_back(); break;
      case Stmt::Expr:
      case Stmt::Not: {
        Opt x;
        value_use(st.a, scope, x.need, x.out);
        if (st.k == Stmt::Not) {
          for (int v : x.out) {
            const std::string& n = names[v];
            if (!(n.size() > 1 && n[0] == '$' && n[1] == 'w')) x.need.insert(v);   // (a wildcard index is local to the negation)
          }
          x.out.clear();
        }
        o.push_back(std::move(x));
        break;
      }
      case Stmt::Assign:
      case Stmt::Unify: {
        for (int side = 0; side < (st.k == Stmt::Unify ? 2 : 1); ++side) {
          Opt x;
          value_use(side ? st.a : st.b, scope, x.need, x.out);
          pattern_use(side ? st.b : st.a, scope, x.need, x.out);
          o.push_back(std::move(x));
        }
        break;
      }
      case Stmt::SomeIn: {
        Opt x;
        value_use(st.c, scope, x.need, x.out);
        pattern_use(st.a, scope, x.need, x.out);
        pattern_use(st.b, scope, x.need, x.out);
        o.push_back(std::move(x));
        break;
      }
      default: o.emplace_back(); break;
    }
    return o;
  }
  static bool subset(const Set& a, const Set& b) {
    for (int v : a)
      if (!b.count(v)) return false;
    return true;
  }
  std::vector<Stmt> reorder(const std::vector<Stmt>& body0, const Set& bound, Set scope) {
    for (auto& st : body0) all_vars(st, scope, false);   // closure candidates of nested comprehensions: everything named so far
    std::vector<Stmt> body;
    for (auto& st : body0) body.push_back(nested(st, scope));
    std::vector<std::vector<Opt>> opts;
    for (auto& st : body) opts.push_back(options(st, scope));
    auto runnable = [&](size_t i, const Set& b) -> const Opt* {
      for (auto& o : opts[i])
        if (subset(o.need, b)) return &o;
      return nullptr;
    };
    {
      Set b = bound;
      bool ok = true;
      for (size_t i = 0; i < body.size() && ok; ++i) {
        const Opt* o = runnable(i, b);
        if (!o) ok = false;
        else b.insert(o->out.begin(), o->out.end());
      }
      if (ok) return body;
    }
    Set b = bound;
    std::vector<size_t> left(body.size()), order;
    for (size_t i = 0; i < left.size(); ++i) left[i] = i;
    while (!left.empty()) {
      bool moved = false;
      for (size_t j = 0; j < left.size(); ++j) {
        if (const Opt* o = runnable(left[j], b)) {
          b.insert(o->out.begin(), o->out.end());
          order.push_back(left[j]);
          left.erase(left.begin() + (long)j);
          moved = true;
          break;
        }
      }
      if (!moved) {   // nothing can run: keep what is left as written (the evaluator reports the unsafe variable)
        order.insert(order.end(), left.begin(), left.end());
        break;
      }
    }
    std::vector<Stmt> out;
    for (size_t i : order) out.push_back(body[i]);
    return out;
  }
  // comprehension bodies inside a term / statement
  TP nested(const TP& t, const Set& scope) {
    if (!t || t->k == TK::Scalar || t->k == TK::Var) return t;
    auto x = std::make_shared<Term>(*t);
    x->head = nested(t->head, scope);
    for (auto& a : x->args) a = nested(a, scope);
    for (auto& kv : x->kvs) kv = {nested(kv.first, scope), nested(kv.second, scope)};
    if (t->k == TK::ArrCompr || t->k == TK::SetCompr || t->k == TK::ObjCompr) {
      x->body = reorder(t->body, scope, scope);
      Set inner = scope;
      for (auto& st : x->body) all_vars(st, inner, false);
      x->key = nested(t->key, inner);
      x->value = nested(t->value, inner);
    } else {
      x->key = nested(t->key, scope);
      x->value = nested(t->value, scope);
    }
    return x;
  }
  Stmt nested(const Stmt& s, const Set& scope) {
    Stmt o = s;
    o.a = nested(s.a, scope), o.b = nested(s.b, scope), o.c = nested(s.c, scope);
    return o;
  }
  void run() {
    for (auto& kv : m.rules)
      for (auto& r : kv.second) {
        Set bound;
        for (auto& a : r.args) all_vars(a, bound);
        r.body = reorder(r.body, bound, bound);
        for (auto& e : r.els) e.second = reorder(e.second, bound, bound);
        Set inner = bound;
        for (auto& st : r.body) all_vars(st, inner, false);
        r.key = nested(r.key, inner);
        r.value = nested(r.value, inner);
      }
  }
};

// The one compile-time check the reference's tests pin (pkg/gator/fixtures/fixtures.go TemplateCompileError,
// a body that is just the undeclared identifier `f`): a bare variable statement must be bound earlier.
void check_unsafe(const Module& m) {
  for (auto& kv : m.rules)
    for (auto& r : kv.second) {
      std::vector<int> bound;
      for (auto& a : r.args) collect_vars(a, bound);
      for (auto& s : r.body) {
        if (s.k == Stmt::Expr && s.a->k == TK::Var) {
          const std::string& n = s.a->name;
          bool ok = n == "input" || n == "data" || m.is_rule(n) || n[0] == '$';
          for (int v : bound) ok = ok || v == s.a->vid;
          if (!ok) throw RegoError{"rego_unsafe_var_error: var " + n + " is unsafe (line " + std::to_string(s.line) + ")"};
        }
        if (s.a) collect_vars(s.a, bound);
        if (s.b) collect_vars(s.b, bound);
        if (s.c) collect_vars(s.c, bound);
      }
    }
}

}  // namespace

int Module::intern(const std::string& n) {
  auto it = symtab.find(n);
  if (it != symtab.end()) return it->second;
  int id = next_vid++;
  symtab.emplace(n, id);
  return id;
}

// ---- purity: which rules never look at `input` / `data`
static void term_deps(const Module& m, const Term& t, bool& touches_doc, std::vector<std::string>& refs);
static void stmts_deps(const Module& m, const std::vector<Stmt>& b, bool& touches_doc, std::vector<std::string>& refs) {
  for (auto& st : b)
    for (const TP* x : {&st.a, &st.b, &st.c})
      if (*x) term_deps(m, **x, touches_doc, refs);
}
static thread_local bool* g_touches_params = nullptr;   // optional second flag: input.parameters / bare input / data
static void term_deps(const Module& m, const Term& t, bool& touches_doc, std::vector<std::string>& refs) {
  if (t.k == TK::Var) {
    if (t.vid == m.vid_input || t.vid == m.vid_data) {
      touches_doc = true;
      if (g_touches_params) *g_touches_params = true;    // a bare `input` / `data` (reference heads are handled below)
    }
    if (m.is_rule(t.name)) refs.push_back(t.name);
  }
  if (t.k == TK::Ref && t.head && t.head->k == TK::Var && t.head->vid == m.vid_input && !t.args.empty() && t.args[0]->k == TK::Scalar &&
      t.args[0]->val->t == VT::Str && t.args[0]->val->s != "parameters") {
    // input.review... : touches the document but not the parameters
    touches_doc = true;
    for (auto& a : t.args) term_deps(m, *a, touches_doc, refs);
    return;
  }
  if (t.k == TK::Call && m.is_rule(t.name)) refs.push_back(t.name);
  if (t.head) term_deps(m, *t.head, touches_doc, refs);
  for (auto& a : t.args) term_deps(m, *a, touches_doc, refs);
  for (auto& kv : t.kvs) term_deps(m, *kv.first, touches_doc, refs), term_deps(m, *kv.second, touches_doc, refs);
  if (t.key) term_deps(m, *t.key, touches_doc, refs);
  if (t.value) term_deps(m, *t.value, touches_doc, refs);
  stmts_deps(m, t.body, touches_doc, refs);
}
static void compute_purity(Module& m) {
  std::map<std::string, std::vector<std::string>> refs;
  std::map<std::string, bool> impure, uses_params;
  for (auto& kv : m.rules) {
    bool doc = false, par = false;
    g_touches_params = &par;
    auto& rf = refs[kv.first];
    for (auto& r : kv.second) {
      for (auto& a : r.args) term_deps(m, *a, doc, rf);
      if (r.key) term_deps(m, *r.key, doc, rf);
      if (r.value) term_deps(m, *r.value, doc, rf);
      stmts_deps(m, r.body, doc, rf);
      for (auto& e : r.els) {
        if (e.first) term_deps(m, *e.first, doc, rf);
        stmts_deps(m, e.second, doc, rf);
      }
    }
    impure[kv.first] = doc;
    uses_params[kv.first] = par;
    g_touches_params = nullptr;
  }
  for (bool changed = true; changed;) {
    changed = false;
    for (auto& kv : refs)
      if (!impure[kv.first])
        for (auto& r : kv.second)
          if (impure[r]) {
            impure[kv.first] = true;
            changed = true;
            break;
          }
  }
  for (bool changed = true; changed;) {
    changed = false;
    for (auto& kv : refs)
      if (!uses_params[kv.first])
        for (auto& r : kv.second)
          if (uses_params[r]) {
            uses_params[kv.first] = true;
            changed = true;
            break;
          }
  }
  for (auto& kv : m.rules) {
    if (kv.second[0].kind == Rule::Func) m.pure_fn[kv.first] = !impure[kv.first];
    m.param_free[kv.first] = !uses_params[kv.first];
  }
}

std::shared_ptr<Module> rego_parse(const std::string& src, const std::vector<std::string>& libs) {
  static std::atomic<uint64_t> uid{1};
  auto m = std::make_shared<Module>();
  m->uid = uid++;
  m->vid_input = m->intern("input");
  m->vid_data = m->intern("data");
  // libs first (pre-scan each for its package and rule names, then parse it into `m` under its package prefix)
  std::set<std::string> lib_pkgs;
  std::vector<std::set<std::string>> lib_rules(libs.size());
  std::vector<std::string> lib_pkg(libs.size());
  for (size_t i = 0; i < libs.size(); ++i) {
    Module scratch;
    scratch.vid_input = scratch.intern("input");
    scratch.vid_data = scratch.intern("data");
    Parser p0(libs[i], scratch);
    p0.lenient_imports = true;
    p0.parse_module();
    // frameworks' regorewriter: a lib lives under data.lib (constraint/pkg/regorewriter, libs must be `package lib.<...>`)
    if (scratch.package != "lib" && scratch.package.rfind("lib.", 0) != 0)
      throw RegoError{"rego_compile_error: lib package `" + scratch.package + "` must begin with `lib`"};
    lib_pkg[i] = "data." + scratch.package;
    lib_pkgs.insert(lib_pkg[i]);
    for (auto& kv : scratch.rules) lib_rules[i].insert(kv.first);
  }
  for (size_t i = 0; i < libs.size(); ++i) {
    Parser pl(libs[i], *m);
    pl.prefix = lib_pkg[i] + ".";
    pl.own_rules = &lib_rules[i];
    pl.lib_pkgs = &lib_pkgs;
    pl.parse_module();
  }
  Parser p(src, *m);
  p.lib_pkgs = &lib_pkgs;
  p.parse_module();
  for (auto& kv : m->rules) {
    Rule::Kind k0 = kv.second[0].kind;
    for (auto& r : kv.second)
      if ((r.kind == Rule::Func) != (k0 == Rule::Func))
        throw RegoError{"rego_type_error: conflicting rules named " + kv.first};
  }
  EveryDesugar{*m}.run();
  ComprScoper{*m}.run();
  BodyOrder{*m, {}}.run();
  check_unsafe(*m);
  compute_purity(*m);
  return m;
}

// ------------------------------------------------------------------------------------- printing
static void stmt_str(const Stmt& s, std::string& out);
static void term_rec(const Term& t, std::string& out) {
  switch (t.k) {
    case TK::Scalar: out += fmt_value(t.val, false); break;
    case TK::Var: out += (t.name.size() > 1 && t.name[0] == '$' && t.name[1] == 'w') ? std::string("_") : t.name; break;   // every wildcard is a fresh variable
    case TK::Ref:
      term_rec(*t.head, out);
      for (auto& a : t.args) {
        out.push_back('[');
        term_rec(*a, out);
        out.push_back(']');
      }
      break;
    case TK::Call:
      out += t.name;
      out.push_back('(');
      for (size_t i = 0; i < t.args.size(); ++i) {
        if (i) out += ", ";
        term_rec(*t.args[i], out);
      }
      out.push_back(')');
      break;
    case TK::Array:
    case TK::Set:
      out += t.k == TK::Array ? "[" : "{";
      for (size_t i = 0; i < t.args.size(); ++i) {
        if (i) out += ", ";
        term_rec(*t.args[i], out);
      }
      if (t.k == TK::Set && t.args.empty()) out += "set()";
      out += t.k == TK::Array ? "]" : "}";
      break;
    case TK::Object:
      out.push_back('{');
      for (size_t i = 0; i < t.kvs.size(); ++i) {
        if (i) out += ", ";
        term_rec(*t.kvs[i].first, out);
        out += ": ";
        term_rec(*t.kvs[i].second, out);
      }
      out.push_back('}');
      break;
    case TK::ArrCompr:
    case TK::SetCompr:
    case TK::ObjCompr:
      out += t.k == TK::ArrCompr ? "[" : "{";
      if (t.k == TK::ObjCompr) {
        term_rec(*t.key, out);
        out += ": ";
      }
      term_rec(*t.value, out);
      out += " | ";
      for (size_t i = 0; i < t.body.size(); ++i) {
        if (i) out += "; ";
        stmt_str(t.body[i], out);
      }
      out += t.k == TK::ArrCompr ? "]" : "}";
      break;
  }
}
static void stmt_str(const Stmt& s, std::string& out) {
  switch (s.k) {
    case Stmt::Expr: term_rec(*s.a, out); break;
    case Stmt::Not:
      out += "not ";
      term_rec(*s.a, out);
      break;
    case Stmt::Assign:
    case Stmt::Unify:
      term_rec(*s.a, out);
      out += s.k == Stmt::Assign ? " := " : " = ";
      term_rec(*s.b, out);
      break;
    case Stmt::Some: out += "some"; break;
    case Stmt::SomeIn:
      out += "some ";
      if (s.a) {
        term_rec(*s.a, out);
        out += ", ";
      }
      term_rec(*s.b, out);
      out += " in ";
      term_rec(*s.c, out);
      break;
  }
}
std::string term_str(const Term& t) {
  std::string out;
  term_rec(t, out);
  return out;
}

}  // namespace gk

namespace gk {
// Canonical text of every definition of a rule (used to recognise identical helper rules of different templates).
std::string rule_str(const Module& m, const std::string& name) {
  std::string out;
  auto it = m.rules.find(name);
  if (it == m.rules.end()) return out;
  for (auto& r : it->second) {
    out += name;
    out += r.kind == Rule::Func ? "(" : r.kind == Rule::PSet ? "[" : r.kind == Rule::PObj ? "{" : "=";
    for (auto& a : r.args) out += term_str(*a) + ",";
    if (r.key) out += term_str(*r.key);
    out += "=";
    if (r.value) out += term_str(*r.value);
    out += r.is_default ? " default{" : " {";
    for (auto& s : r.body) {
      stmt_str(s, out);
      out += ";";
    }
    out += "}";
    for (auto& e : r.els) {
      out += "else=";
      if (e.first) out += term_str(*e.first);
      out += "{";
      for (auto& s : e.second) {
        stmt_str(s, out);
        out += ";";
      }
      out += "}";
    }
    out += "\n";
  }
  return out;
}
}  // namespace gk
