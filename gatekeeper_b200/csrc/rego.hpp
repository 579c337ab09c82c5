// Rego front-end: AST, parser and the concrete (host) evaluator.
//
// The engine never interprets Rego per (constraint, object) pair to DECIDE a violation -- that is the job
// of the lowered predicate program on the GPU (lower.hpp / kernels.cu).  The concrete evaluator is used
//   (1) at AddConstraint time to fold everything that depends only on `input.parameters`,
//   (2) at flatten time to extract parameter-independent feature columns from each object, and
//   (3) after the kernel, to render `msg`/`details` for pairs the GPU already flagged.
// Language subset: what every in-tree ConstraintTemplate uses; anything else is
// rejected at AddTemplate with a rego_* error, like a compile error from the reference's Rego driver.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "val.hpp"

namespace gk {

struct RegoError {
  std::string msg;
};

enum class TK : uint8_t { Scalar, Var, Ref, Call, Array, Object, Set, ArrCompr, SetCompr, ObjCompr };

struct Term;
using TP = std::shared_ptr<const Term>;

struct Stmt {
  enum K : uint8_t { Expr, Not, Assign, Unify, Some, SomeIn, Every } k = Expr;
  TP a, b, c;   // Expr/Not: a ; Assign/Unify: a,b ; SomeIn: a=key(or null), b=value, c=collection
  int line = 0;
  // Every (`every k, v in c { body }`) exists only between the parser and its desugaring pass (rego_parse.cpp): no evaluator sees it
  std::vector<Stmt> body;
};

struct Term {
  TK k = TK::Scalar;
  VP val;                               // Scalar
  int vid = -1;                         // Var: symbol id (module-wide); wildcards get fresh ids
  std::string name;                     // Var / Call name
  TP head;                              // Ref
  std::vector<TP> args;                 // Ref path / Call args / Array+Set items
  std::vector<std::pair<TP, TP>> kvs;   // Object
  TP key, value;                        // comprehension head (ObjCompr: key+value; others: value)
  std::vector<Stmt> body;               // comprehension body
  int line = 0;
  // resolved lazily by the evaluator (same value from every thread): is `name` a rule of the module / a builtin id
  mutable signed char is_rule_ = -1;
  mutable const void* rules_ = nullptr;
};

struct Rule {
  enum Kind : uint8_t { Complete, Func, PSet, PObj } kind = Complete;
  std::string name;
  std::vector<TP> args;
  TP key, value;
  std::vector<Stmt> body;
  bool is_default = false;
  bool has_body = false;
  std::vector<std::pair<TP, std::vector<Stmt>>> els;
  int line = 0;
};

struct Module {
  std::string package;
  std::map<std::string, std::vector<Rule>> rules;
  std::unordered_map<std::string, int> symtab;   // var name -> vid
  int next_vid = 0;
  int vid_input = -1, vid_data = -1;
  uint64_t uid = 0;                               // unique per parsed module (column-key namespace)
  // functions whose value depends on their arguments only (no `input` / `data`, transitively): their results
  // can be memoised across objects
  std::unordered_map<std::string, bool> pure_fn;
  // rules whose value does not depend on `input.parameters` (nor on `data`), transitively: for one object their extents
  // are the same under every constraint of the template
  std::unordered_map<std::string, bool> param_free;
  int intern(const std::string& n);
  bool is_rule(const std::string& n) const { return rules.count(n) != 0; }
};

// `libs`: the template's `spec.targets[].libs` modules (package lib.<...>), importable as data.lib.<...>
std::shared_ptr<Module> rego_parse(const std::string& src, const std::vector<std::string>& libs = {});   // throws RegoError
std::string term_str(const Term& t);                          // debug / canonical printing
std::string rule_str(const Module& m, const std::string& name); // canonical text of all definitions of a rule

// ---- concrete evaluator ---------------------------------------------------------------------------
struct Env {
  std::vector<std::pair<int, VP>> b;
  const VP* find(int vid) const {
    for (size_t i = b.size(); i-- > 0;)
      if (b[i].first == vid) return &b[i].second;
    return nullptr;
  }
  size_t mark() const { return b.size(); }
  void undo(size_t m) { b.resize(m); }
  void bind(int vid, VP v) { b.emplace_back(vid, std::move(v)); }
};

// Non-owning callable reference: continuations are only ever passed DOWN the evaluator's recursion and never stored,
// so they need neither a copy nor a heap allocation (std::function allocates for every capture list > 16 bytes).
template <class Sig>
class FnRef;
template <class R, class... A>
class FnRef<R(A...)> {
  void* obj_;
  R (*call_)(void*, A...);

 public:
  template <class F, class = typename std::enable_if<!std::is_same<typename std::decay<F>::type, FnRef>::value>::type>
  FnRef(F&& f)
      : obj_((void*)std::addressof(f)),
        call_([](void* o, A... a) -> R { return (*static_cast<typename std::remove_reference<F>::type*>(o))(std::forward<A>(a)...); }) {}
  R operator()(A... a) const { return call_(obj_, std::forward<A>(a)...); }
};

// Continuations return true to STOP the search (first-solution / negation probes).
using ValK = FnRef<bool(const VP&)>;
using EnvK = FnRef<bool()>;

class Eval {
 public:
  Eval(const Module& m, VP input, VP data = nullptr) : m_(m), input_(std::move(input)), data_(std::move(data)) {}
  void reset_input(VP input) {
    input_ = std::move(input);
    cache_.clear();
    cache_has_.clear();
  }
  // same review, other parameters: extents of parameter-free rules stay valid
  void reset_parameters(VP input) {
    input_ = std::move(input);
    for (auto it = cache_has_.begin(); it != cache_has_.end();) {
      auto pf = m_.param_free.find(it->first);
      if (pf != m_.param_free.end() && pf->second) {
        ++it;
      } else {
        cache_.erase(it->first);
        it = cache_has_.erase(it);
      }
    }
  }
  bool eval_body(const std::vector<Stmt>& body, size_t i, Env& env, const EnvK& k);
  bool eval_term(const TP& t, Env& env, const ValK& k);
  VP eval_first(const TP& t, Env& env);                 // first solution or nullptr (undefined)
  VP rule_value(const std::string& name);               // nullptr if undefined
  VP call_function(const std::string& name, const std::vector<VP>& args);   // nullptr if undefined
  VP call_function(const std::vector<Rule>& rules, const std::vector<VP>& args);
  // the definitions of the function a Call term names (resolved once per term), or nullptr for a builtin / non-function
  const std::vector<Rule>* function_rules(const Term& t, bool* names_other_rule = nullptr) const;
  bool unify_val(const TP& pat, const VP& val, Env& env, const EnvK& k);
  bool is_ground(const TP& t, const Env& env) const;
  const Module& module() const { return m_; }

 private:
  bool eval_stmt(const Stmt& st, Env& env, const EnvK& k);
  bool unify(const TP& a, const TP& b, Env& env, const EnvK& k);
  bool walk(const VP& cur, const std::vector<TP>& path, size_t i, Env& env, const ValK& k);
  bool eval_call(const Term& t, Env& env, const ValK& k);
  bool eval_seq(const std::vector<TP>& items, size_t i, std::vector<VP>& acc, Env& env, const EnvK& k);
  VP rule_chain(const Rule& r, Env& env);
  bool var_unbound(const Term& t, const Env& env) const;
  const Module& m_;
  VP input_, data_;
  std::unordered_map<std::string, VP> cache_;   // rule extents; nullptr entries = undefined
  std::unordered_map<std::string, bool> cache_has_;
  // pure functions on scalar arguments (survives reset_input): per function, argument tuple -> value
  std::unordered_map<const void*, std::unordered_map<std::string, VP>> fn_memo_;
  std::unordered_map<const void*, bool> pure_of_;
  int depth_ = 0;
};

// builtins: returns nullptr for undefined (type errors are undefined, as in OPA's non-strict mode).
// `known` is set to false when the name is not a builtin at all.
VP call_builtin(const std::string& name, const std::vector<VP>& args, bool* known);
bool is_builtin(const std::string& name);

}  // namespace gk
