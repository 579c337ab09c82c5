// TEST / BENCH INFRASTRUCTURE -- not part of the product library.
//
// CPU restatement, in C++, of the reference's Client.Review loop for the audit shape: for every object, for every constraint
// that applies at the enforcement point, run the spec.match pre-filter (pkg/mutation/match/match.go:32-268, restated from
// oracle/k8s.py function by function) and, on a match, evaluate the template's `violation` rule with that constraint's
// parameters using the repo's concrete Rego evaluator (csrc/rego_eval.cpp -- the stand-in for OPA's topdown, which is not in
// the reference tree).  One JSON parse per object, one evaluation per matching (constraint, object) pair, std::thread over
// the host CPUs.  This is what `bench.py --impl reference` times (kind "cpp-restatement"); tests/test_cpu_ref.py checks it
// against the Python oracle on the same objects.  Built by gatekeeper_b200/build.py into oracle/_build/libgk_cpuref.so from
// the host objects of the engine (no CUDA, no backend).
#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../gatekeeper_b200/csrc/engine.hpp"

using namespace gk;

namespace {

std::string sfield(const VP& o, const char* k) {
  VP v = obj_get(o, k);
  return v && v->t == VT::Str ? v->s : std::string();
}

// Wildcard.Matches / MatchesGenerateName -- pkg/wildcard/wildcard.go:17-41
bool wild(const std::string& w, const std::string& s) {
  const bool pre = !w.empty() && w.front() == '*', suf = w.size() > (pre ? 1u : 0u) && w.back() == '*';
  if (w == "*") return true;
  const std::string core = w.substr(pre ? 1 : 0, w.size() - (pre ? 1 : 0) - (suf ? 1 : 0));
  if (pre && suf) return s.find(core) != std::string::npos;
  if (pre) return s.size() >= core.size() && s.compare(s.size() - core.size(), core.size(), core) == 0;
  if (suf) return s.compare(0, core.size(), core) == 0;
  return s == w;
}
bool wild_gen(const std::string& w, const std::string& s) {
  const bool pre = !w.empty() && w.front() == '*', suf = w.size() > (pre ? 1u : 0u) && w.back() == '*';
  if (w == "*") return true;
  const std::string core = w.substr(pre ? 1 : 0, w.size() - (pre ? 1 : 0) - (suf ? 1 : 0));
  if (pre && suf) return s.find(core) != std::string::npos;
  if (suf) return s.compare(0, core.size(), core) == 0;
  return false;
}

const Node* labels(const VP& o) {
  VP md = obj_get(o, "metadata");
  if (!md || md->t != VT::Obj) return nullptr;
  VP l = obj_get(md, "labels");
  return l && l->t == VT::Obj ? l.get() : nullptr;
}
bool label_get(const Node* ls, const std::string& key, std::string* val) {
  if (!ls) return false;
  for (auto& e : ls->kv)
    if (e.first->s == key) {
      *val = e.second->t == VT::Str ? e.second->s : std::string("\x01non-string");
      return true;
    }
  return false;
}

struct MatchErr {
  std::string msg;
};

// LabelSelectorAsSelector + Selector.Matches (apimachinery): matchLabels{k:v} == In(k,[v]); requirements ANDed
bool selector(const VP& sel, const Node* ls) {
  VP ml = obj_get(sel, "matchLabels");
  if (ml && ml->t == VT::Obj)
    for (auto& e : ml->kv) {
      std::string v;
      if (!label_get(ls, e.first->s, &v) || e.second->t != VT::Str || v != e.second->s) return false;
    }
  VP me = obj_get(sel, "matchExpressions");
  if (me && me->t == VT::Arr)
    for (auto& r : me->items) {
      const std::string key = sfield(r, "key"), op = sfield(r, "operator");
      VP vals = obj_get(r, "values");
      const size_t nv = vals && vals->t == VT::Arr ? vals->items.size() : 0;
      std::string v;
      const bool has = label_get(ls, key, &v);
      bool in = false;
      for (size_t j = 0; j < nv && has; ++j) in = in || (vals->items[j]->t == VT::Str && vals->items[j]->s == v);
      if (op == "In") {
        if (nv == 0) throw MatchErr{"invalid label selector"};
        if (!in) return false;
      } else if (op == "NotIn") {
        if (nv == 0) throw MatchErr{"invalid label selector"};
        if (has && in) return false;
      } else if (op == "Exists") {
        if (nv) throw MatchErr{"invalid label selector"};
        if (!has) return false;
      } else if (op == "DoesNotExist") {
        if (nv) throw MatchErr{"invalid label selector"};
        if (has) return false;
      } else {
        throw MatchErr{"invalid label selector operator"};
      }
    }
  return true;
}

bool in_list(const VP& list, const std::string& s) {
  if (!list || list->t != VT::Arr) return false;
  for (auto& x : list->items)
    if (x->t == VT::Str && (x->s == s)) return true;
  return false;
}
bool has_star(const VP& list) { return in_list(list, "*"); }
size_t len_of(const VP& list) { return list && list->t == VT::Arr ? list->items.size() : 0; }

// match.Matches: the 8 matchers in their fixed order, early exit -- match.go:41-62
bool matches(const VP& m, const VP& obj, const VP& ns, const std::string& src) {
  std::string g, v, k;
  split_gv(obj, g, v, k);
  const std::string name = meta_str(obj, "name"), objns = meta_str(obj, "namespace");
  const bool is_ns = k == "Namespace" && g.empty();
  // 1 kinds -- :181-201
  VP kinds = obj_get(m, "kinds");
  if (len_of(kinds)) {
    bool any = false;
    for (auto& kk : kinds->items) {
      VP ks = obj_get(kk, "kinds"), gs = obj_get(kk, "apiGroups");
      if (!(len_of(ks) == 0 || has_star(ks) || in_list(ks, k))) continue;
      if (len_of(gs) == 0 || has_star(gs) || in_list(gs, g)) {
        any = true;
        break;
      }
    }
    if (!any) return false;
  }
  // 2 scope -- :214-227
  const bool has_ns = !objns.empty() || ns;
  const std::string scope = sfield(m, "scope");
  if (scope == "Cluster" && !(is_ns || !has_ns)) return false;
  if (scope == "Namespaced" && !(!is_ns && has_ns)) return false;
  // 3 / 4 namespaces, excludedNamespaces -- :118-179
  auto ns_name = [&](std::string* out) {
    if (is_ns) return *out = name, true;
    if (ns) return *out = meta_str(ns, "name"), true;
    if (!objns.empty()) return *out = objns, true;
    return false;
  };
  VP nss = obj_get(m, "namespaces"), ex = obj_get(m, "excludedNamespaces");
  std::string nn;
  if (len_of(nss) && ns_name(&nn)) {
    bool any = false;
    for (auto& p : nss->items) any = any || (p->t == VT::Str && wild(p->s, nn));
    if (!any) return false;
  }
  if (len_of(ex) && ns_name(&nn))
    for (auto& p : ex->items)
      if (p->t == VT::Str && wild(p->s, nn)) return false;
  // 5 labelSelector -- :103-116
  VP ls = obj_get(m, "labelSelector");
  if (ls && ls->t != VT::Null && !selector(ls, labels(obj))) return false;
  // 6 namespaceSelector -- :73-101
  VP nsel = obj_get(m, "namespaceSelector");
  if (nsel && nsel->t != VT::Null && !(!is_ns && !ns && objns.empty())) {
    if (is_ns) {
      if (!selector(nsel, labels(obj))) return false;
    } else {
      if (!ns) throw MatchErr{"namespace selector for namespace-scoped object but missing Namespace"};
      if (!selector(nsel, labels(ns))) return false;
    }
  }
  // 7 name -- :203-212
  const std::string mname = sfield(m, "name");
  if (!mname.empty() && !(wild(mname, name) || wild_gen(mname, meta_str(obj, "generateName")))) return false;
  // 8 source -- :229-253
  std::string msrc = sfield(m, "source");
  if (msrc.empty()) msrc = "All";
  else if (msrc != "All" && msrc != "Generated" && msrc != "Original") throw MatchErr{"invalid source field"};
  if (src.empty() && msrc != "All") throw MatchErr{"source field not specified"};
  if (msrc == "All") return true;
  if (src != "All" && src != "Generated" && src != "Original") throw MatchErr{"invalid source field"};
  return msrc == src;
}

struct CpuRef {
  Engine eng{1};
  std::vector<std::string> keys;
};

char* dup_str(const std::string& s) {
  char* p = (char*)malloc(s.size() + 1);
  if (p) memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

}  // namespace

extern "C" {

void* gk_cpuref_create(void) { return new CpuRef(); }
void gk_cpuref_destroy(void* h) { delete static_cast<CpuRef*>(h); }

int gk_cpuref_add_template(void* h, const char* kind, const char* rego, size_t len, char** err) {
  try {
    static_cast<CpuRef*>(h)->eng.add_template(kind, std::string(rego, len));
    return 0;
  } catch (RegoError& e) {
    if (err) *err = dup_str(e.msg);
    return -1;
  }
}
int gk_cpuref_add_constraint(void* h, const char* json, size_t len, char** err) {
  try {
    static_cast<CpuRef*>(h)->eng.add_constraint(std::string(json, len));
    return 0;
  } catch (RegoError& e) {
    if (err) *err = dup_str(e.msg);
    return -1;
  } catch (JsonError& e) {
    if (err) *err = dup_str(e.msg);
    return -1;
  }
}
int gk_cpuref_put_namespace(void* h, const char* name, const char* json, size_t len) {
  try {
    static_cast<CpuRef*>(h)->eng.put_namespace(name, std::string(json, len));
    return 0;
  } catch (...) {
    return -1;
  }
}
uint32_t gk_cpuref_constraint_count(void* h) {
  auto* r = static_cast<CpuRef*>(h);
  auto c = r->eng.compiled();
  r->keys.clear();
  for (auto* k : c->order) r->keys.push_back(k->kind + "/" + k->name);
  return (uint32_t)r->keys.size();
}
const char* gk_cpuref_constraint_key(void* h, uint32_t i) {
  auto* r = static_cast<CpuRef*>(h);
  return i < r->keys.size() ? r->keys[i].c_str() : nullptr;
}

// Reviews `n` plain objects at enforcement point `ep` on `threads` host threads.  pair_totals[c] = objects violating constraint
// c (in gk_cpuref_constraint_key order), *n_results = distinct {msg, details} results, *n_errors = matcher errors (autorejects),
// *seconds = wall time of the review loop.
int gk_cpuref_review_blob(void* h, const char* buf, const uint64_t* off, size_t n, const char* source, const char* ep, int threads,
                          uint64_t* pair_totals, uint64_t* n_results, uint64_t* n_errors, double* seconds, char** err) {
  auto* r = static_cast<CpuRef*>(h);
  try {
    auto c = r->eng.compiled();
    const size_t C = c->order.size();
    std::vector<uint32_t> active;
    r->eng.active_mask(*c, ep ? ep : "", active);
    const std::string src = source ? source : "";
    const uint8_t src_code = src == "Original" ? GK_SRC_ORIGINAL : src == "Generated" ? GK_SRC_GENERATED : src == "All" ? GK_SRC_ALL : src.empty() ? GK_SRC_EMPTY : GK_SRC_INVALID;
    const size_t T = (size_t)std::max(1, threads);
    std::vector<std::vector<uint64_t>> tot(T, std::vector<uint64_t>(C, 0));
    std::vector<uint64_t> res(T, 0), errs(T, 0);
    std::vector<std::string> fail(T);
    std::atomic<size_t> next{0};
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](size_t t) {
      try {
        std::unordered_map<const Module*, std::unique_ptr<Eval>> evals;
        std::vector<VP> params(C);
        for (size_t ci = 0; ci < C; ++ci) params[ci] = v_deep_copy(c->order[ci]->params);   // thread-private: no shared reference counts
        const std::map<std::string, VP> ns_cache = r->eng.namespaces_snapshot();
        for (;;) {
          const size_t lo = next.fetch_add(256);
          if (lo >= n) break;
          for (size_t i = lo; i < std::min(n, lo + 256); ++i) {
            ObjIn in;
            in.json = buf + off[i];
            in.len = (size_t)(off[i + 1] - off[i]);
            in.source = src_code;
            std::string e;
            VP obj, old, ns;
            VP doc = r->eng.review_doc(in, &obj, &old, &ns, &e, &ns_cache);   // one parse per object (the reference: one per constraint)
            if (!doc || !obj) continue;
            bool first = true;
            for (size_t ci = 0; ci < C; ++ci) {
              if (!active[ci]) continue;
              const Constraint& con = *c->order[ci];
              bool m = true;
              if (con.match.has) {
                try {
                  m = matches(con.match.raw, obj, ns, src);
                } catch (MatchErr&) {
                  ++errs[t];
                  continue;
                }
              }
              if (!m) continue;
              const Module& mod = *c->mods[ci];
              VP inp = v_obj({{v_str("review"), doc}, {v_str("parameters"), params[ci]}});
              auto& slot = evals[&mod];
              if (!slot) slot.reset(new Eval(mod, inp));
              else slot->reset_input(inp);   // a fresh query per (constraint, object) pair, like the reference's driver
              (void)first;
              VP vs = slot->rule_value("violation");
              if (vs && !vs->items.empty()) {
                ++tot[t][ci];
                res[t] += vs->items.size();
              }
            }
            for (auto& ev : evals) ev.second->reset_input(nullptr);
          }
        }
      } catch (RegoError& x) {
        fail[t] = x.msg;
      } catch (std::exception& x) {
        fail[t] = x.what();
      }
    };
    std::vector<std::thread> th;
    for (size_t t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& f : fail)
      if (!f.empty()) throw RegoError{f};
    for (size_t ci = 0; ci < C && pair_totals; ++ci) {
      pair_totals[ci] = 0;
      for (size_t t = 0; t < T; ++t) pair_totals[ci] += tot[t][ci];
    }
    if (n_results) {
      *n_results = 0;
      for (auto v : res) *n_results += v;
    }
    if (n_errors) {
      *n_errors = 0;
      for (auto v : errs) *n_errors += v;
    }
    return 0;
  } catch (RegoError& e) {
    if (err) *err = dup_str(e.msg);
    return -1;
  }
}

void gk_cpuref_free_str(char* s) { free(s); }

}  // extern "C"
