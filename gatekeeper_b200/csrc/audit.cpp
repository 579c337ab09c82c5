// Audit-sweep aggregation and admission message assembly: the host work the reference does around Client.Review.
//
//   AuditRun            addAuditResponsesToUpdateLists   pkg/audit/manager.go:886-945
//   LimitQueue / sv_less LimitQueue, SVQueue.Less        pkg/audit/manager.go:117-202
//   truncate_string     truncateString                   pkg/audit/manager.go:1043-1052
//   report()            updateConstraintStatus           pkg/audit/manager.go:984-1041 (pops the heap: descending order)
//   validation_messages getValidationMessages            pkg/webhook/policy.go:238-355
//
// The kernel has already decided which (constraint, object) pairs violate; this code renders the messages of the flagged
// pairs (all host cores), counts results and keeps the K smallest per constraint.
#include "audit.hpp"

#include <algorithm>
#include <thread>

#include "val.hpp"

namespace gk {

std::string truncate_string(const std::string& s, size_t size) {
  if (s.size() <= size) return s;
  if (size > 3) size -= 3;
  return s.substr(0, size) + "...";
}

// SVQueue.Less is "greater than" so that Go's min-heap pops the LARGEST; sv_less is the natural order underneath it.
bool sv_less(const StatusViolation& a, const StatusViolation& b) {
  if (a.group != b.group) return a.group < b.group;
  if (a.version != b.version) return a.version < b.version;
  if (a.kind != b.kind) return a.kind < b.kind;
  if (a.ns != b.ns) return a.ns < b.ns;
  if (a.name != b.name) return a.name < b.name;
  if (a.message != b.message) return a.message < b.message;
  return a.action < b.action;
}

void LimitQueue::push(StatusViolation v) {
  if (limit == 0) return;
  if (heap.size() >= limit && !sv_less(v, heap.front())) return;   // would be popped straight away
  heap.push_back(std::move(v));
  std::push_heap(heap.begin(), heap.end(), sv_less);
  while (heap.size() > limit) {
    std::pop_heap(heap.begin(), heap.end(), sv_less);
    heap.pop_back();
  }
}

std::vector<StatusViolation> LimitQueue::drain_descending() {
  std::vector<StatusViolation> out;
  while (!heap.empty()) {
    std::pop_heap(heap.begin(), heap.end(), sv_less);
    out.push_back(std::move(heap.back()));
    heap.pop_back();
  }
  return out;
}

void AuditRun::fold(const std::string& key, StatusViolation sv) {
  auto& pc = per_constraint[key];
  if (pc.queue.limit != limit) pc.queue.limit = limit;
  pc.total++;
  by_action[sv.action]++;
  sv.message = truncate_string(sv.message, msg_size);
  pc.queue.push(std::move(sv));
}

void AuditRun::merge(AuditRun& o) {
  for (auto& kv : o.per_constraint) {
    auto& pc = per_constraint[kv.first];
    pc.queue.limit = limit;
    pc.total += kv.second.total;
    for (auto& sv : kv.second.queue.heap) pc.queue.push(std::move(sv));
  }
  for (auto& kv : o.by_action) by_action[kv.first] += kv.second;
  objects += o.objects;
  results += o.results;
  rendered_pairs += o.rendered_pairs;
  counted_pairs += o.counted_pairs;
}

namespace {
// (namespace, name) of object o as the reference orders results (SVQueue.Less after group / version / kind): a Namespace's own
// namespace is empty; the header's nsname is the object's metadata.namespace otherwise
struct IdView {
  const BatchIdentity& id;
  void get(uint32_t o, const uint8_t*& ns, uint32_t& nsl, const uint8_t*& nm, uint32_t& nml) const {
    const uint32_t fl = id.flags[o];
    if ((fl & GK_F_IS_NS) || !(fl & GK_F_NSNAME)) ns = nullptr, nsl = 0;
    else ns = id.ns_bytes.data() + id.ns_off[o], nsl = id.ns_off[o + 1] - id.ns_off[o];
    nm = id.name_bytes.data() + id.name_off[o];
    nml = id.name_off[o + 1] - id.name_off[o];
  }
  static int cmp_bytes(const uint8_t* a, uint32_t al, const uint8_t* b, uint32_t bl) {
    const int r = memcmp(a, b, std::min(al, bl));
    return r ? r : (al < bl ? -1 : al > bl ? 1 : 0);
  }
  int cmp(uint32_t a, uint32_t b) const {
    const uint8_t *ans, *anm, *bns, *bnm;
    uint32_t ansl, anml, bnsl, bnml;
    get(a, ans, ansl, anm, anml);
    get(b, bns, bnsl, bnm, bnml);
    const int r = cmp_bytes(ans, ansl, bns, bnsl);
    return r ? r : cmp_bytes(anm, anml, bnm, bnml);
  }
};
}  // namespace

void AuditRun::add_batch(Engine& eng, const Compiled& c, const std::vector<ObjIn>& objs, const uint32_t* viol, const uint32_t* err,
                         uint32_t words, const std::vector<uint32_t>& errlist, const std::string& ep, const BatchIdentity* id, const uint32_t* amb) {
  const size_t n = objs.size();
  const uint32_t C = (uint32_t)c.order.size();
  // ---- lazy path: which pairs need the host at all
  // cand[c] = the objects whose results can still enter constraint c's list.  Constraints that may have several results per pair
  // are evaluated for every pair (their totals count results: pkg/audit/manager.go:886-945).
  const bool lazy = id && id->uniform_gvk && id->flags.size() == n &&
                    (amb || std::any_of(c.single_result.begin(), c.single_result.end(), [](uint8_t x) { return x != 0; }));
  // exactly one result for a violating pair: proven for the constraint when it was lowered, or for this pair by the ambiguity bitmap
  auto one_result = [&](size_t o, uint32_t cix) { return c.single_result[cix] || (amb && !(amb[o * words + (cix >> 5)] >> (cix & 31u) & 1u)); };
  std::vector<uint32_t> thr_obj(C, 0);     // per single-result constraint: an object holding the limit-th smallest identity
  std::vector<uint8_t> thr_all(C, 1);      // fewer than `limit` flagged objects: all are candidates
  if (lazy && limit) {
    const IdView iv{*id};
    const size_t TT = std::min<size_t>((size_t)std::max(1, eng.threads()), std::max<size_t>(1, n / 4096));
    // per thread, per constraint: max-heap of the `limit` smallest objects seen
    std::vector<std::vector<std::vector<uint32_t>>> heaps(TT, std::vector<std::vector<uint32_t>>(C));
    auto less = [&](uint32_t a, uint32_t b) { return iv.cmp(a, b) < 0; };
    auto scan = [&](size_t t) {
      auto& hp = heaps[t];
      const size_t lo = n * t / TT, hi = n * (t + 1) / TT;
      for (size_t o = lo; o < hi; ++o)
        for (uint32_t w = 0; w < words; ++w) {
          uint32_t vb = viol[o * words + w];
          while (vb) {
            const uint32_t k = (uint32_t)__builtin_ctz(vb);
            vb &= vb - 1;
            const uint32_t cix = w * 32 + k;
            if (cix >= C) continue;
            auto& h = hp[cix];
            if (h.size() < limit) {
              h.push_back((uint32_t)o);
              std::push_heap(h.begin(), h.end(), less);
            } else if (less((uint32_t)o, h.front())) {
              std::pop_heap(h.begin(), h.end(), less);
              h.back() = (uint32_t)o;
              std::push_heap(h.begin(), h.end(), less);
            }
          }
        }
    };
    if (TT == 1) scan(0);
    else {
      std::vector<std::thread> th;
      for (size_t t = 0; t < TT; ++t) th.emplace_back(scan, t);
      for (auto& x : th) x.join();
    }
    for (uint32_t cix = 0; cix < C; ++cix) {
      std::vector<uint32_t> all;
      for (size_t t = 0; t < TT; ++t) all.insert(all.end(), heaps[t][cix].begin(), heaps[t][cix].end());
      if (all.size() < limit) continue;
      std::nth_element(all.begin(), all.begin() + (long)(limit - 1), all.end(), less);
      thr_obj[cix] = all[limit - 1];
      thr_all[cix] = 0;
    }
  }
  std::unordered_map<uint64_t, uint32_t> err_code;
  for (size_t i = 0; i + 2 < errlist.size(); i += 3) err_code[((uint64_t)errlist[i] << 32) | errlist[i + 1]] = errlist[i + 2];
  size_t T = std::min<size_t>((size_t)std::max(1, eng.threads()), std::max<size_t>(1, n / 64));
  std::vector<AuditRun> parts(T);
  std::vector<std::string> errs(T);
  std::atomic<size_t> next{0};
  auto work = [&](size_t t) {
    AuditRun& part = parts[t];
    part.limit = limit;
    part.msg_size = msg_size;
    std::vector<Engine::Flagged> flagged;
    std::vector<Violation> vio;
    Engine::MaterializeCtx mctx;
    std::vector<uint64_t> counted(C, 0);   // pairs of single-result constraints taken from the bitmap
    try {
      for (;;) {
        size_t lo = next.fetch_add(256), hi = std::min(n, lo + 256);
        if (lo >= n) break;
        for (size_t o = lo; o < hi; ++o) {
          flagged.clear();
          for (uint32_t w = 0; w < words; ++w) {
            uint32_t vb = viol[o * words + w], eb = err ? err[o * words + w] : 0u;
            uint32_t any = vb | eb;
            while (any) {
              uint32_t k = (uint32_t)__builtin_ctz(any);
              any &= any - 1;
              uint32_t cix = w * 32 + k;
              if (cix >= C) continue;
              bool is_err = eb >> k & 1u;
              if (lazy && !is_err && one_result(o, cix)) {
                // one result per pair: counted here; evaluated only if the object can still enter the constraint's list
                const bool cand = limit && (thr_all[cix] || IdView{*id}.cmp((uint32_t)o, thr_obj[cix]) <= 0);
                if (!cand) {
                  ++counted[cix];
                  continue;
                }
              }
              part.rendered_pairs++;
              uint32_t code = 0;
              if (is_err) {
                auto it = err_code.find(((uint64_t)o << 32) | c.cons_match[cix]);
                if (it != err_code.end()) code = it->second;
              }
              flagged.push_back({cix, is_err, code});
            }
          }
          if (flagged.empty()) continue;
          vio.clear();
          VP obj;
          eng.materialize_object(c, objs[o], (uint32_t)o, flagged, ep, vio, &obj, &mctx);
          std::string g, v, k, ns, name;
          if (obj) {
            split_gv(obj, g, v, k);
            ns = meta_str(obj, "namespace");
            name = meta_str(obj, "name");
          }
          for (auto& x : vio) {
            const Constraint& con = *c.order[x.constraint];
            StatusViolation sv;
            sv.group = g, sv.version = v, sv.kind = k, sv.ns = ns, sv.name = name;
            sv.message = std::move(x.msg);
            sv.action = x.action;
            sv.scoped_json = std::move(x.scoped_json);
            part.fold(con.kind + "/" + con.name, std::move(sv));
            part.results++;
          }
        }
      }
    } catch (RegoError& e) {
      errs[t] = e.msg;
    } catch (std::exception& e) {
      errs[t] = e.what();
    }
    for (uint32_t cix = 0; cix < C; ++cix)
      if (counted[cix]) {
        const Constraint& con = *c.order[cix];
        auto& pc = part.per_constraint[con.kind + "/" + con.name];
        pc.queue.limit = limit;
        pc.total += counted[cix];
        part.by_action[con.action] += counted[cix];
        part.results += counted[cix];
        part.counted_pairs += counted[cix];
      }
  };
  if (T == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (size_t t = 0; t < T; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  for (auto& e : errs)
    if (!e.empty()) throw RegoError{"audit: " + e};
  for (auto& p : parts) merge(p);
  objects += n;
}

static void sv_json(const StatusViolation& sv, std::string& o) {
  // StatusViolation's JSON tags -- pkg/audit/manager.go:100-109 (namespace and enforcementActions are omitempty)
  o += "{\"group\":";
  json_quote(sv.group, o);
  o += ",\"version\":";
  json_quote(sv.version, o);
  o += ",\"kind\":";
  json_quote(sv.kind, o);
  o += ",\"name\":";
  json_quote(sv.name, o);
  if (!sv.ns.empty()) {
    o += ",\"namespace\":";
    json_quote(sv.ns, o);
  }
  o += ",\"message\":";
  json_quote(sv.message, o);
  o += ",\"enforcementAction\":";
  json_quote(sv.action, o);
  if (!sv.scoped_json.empty() && sv.scoped_json != "[]") o += ",\"enforcementActions\":" + sv.scoped_json;
  o += "}";
}

void AuditRun::add_object_errors(const std::vector<std::string>& errs) {
  for (size_t i = 0; i < errs.size(); ++i)
    if (!errs[i].empty()) {
      if (first_object_errors.size() < 20) first_object_errors.emplace_back(seen_objects + i, errs[i]);
      ++object_errors;
    }
  seen_objects += errs.size();
}

std::string AuditRun::report() {
  std::string o = "{\"objects\":" + std::to_string(objects) + ",\"results\":" + std::to_string(results) + ",\"pairsEvaluated\":" + std::to_string(rendered_pairs) +
                  ",\"pairsCounted\":" + std::to_string(counted_pairs) + ",\"objectErrors\":{\"count\":" +
                  std::to_string(object_errors) + ",\"first\":[";
  for (size_t i = 0; i < first_object_errors.size(); ++i) {
    if (i) o += ",";
    o += "{\"object\":" + std::to_string(first_object_errors[i].first) + ",\"error\":";
    json_quote(first_object_errors[i].second, o);
    o += "}";
  }
  o += "]},\"totalViolations\":{";
  bool first = true;
  for (auto& kv : per_constraint) {
    if (!first) o += ",";
    first = false;
    json_quote(kv.first, o);
    o += ":" + std::to_string(kv.second.total);
  }
  o += "},\"totalViolationsPerEnforcementAction\":{";
  first = true;
  for (auto& kv : by_action) {
    if (!first) o += ",";
    first = false;
    json_quote(kv.first, o);
    o += ":" + std::to_string(kv.second);
  }
  o += "},\"violations\":{";
  first = true;
  for (auto& kv : per_constraint) {
    if (!first) o += ",";
    first = false;
    json_quote(kv.first, o);
    o += ":[";
    LimitQueue q = kv.second.queue;   // reporting does not consume the run
    auto list = q.drain_descending();
    for (size_t i = 0; i < list.size(); ++i) {
      if (i) o += ",";
      sv_json(list[i], o);
    }
    o += "]";
  }
  o += "}}";
  return o;
}

// getValidationMessages -- pkg/webhook/policy.go:238-355: per result, the effective actions are the scoped actions for
// the webhook enforcement point (results with none are dropped) or the constraint's own action; unsupported actions are
// skipped (ValidateEnforcementAction, pkg/util/enforcement_action.go:60-70); "deny" -> denyMsgs, "warn" -> warnMsgs,
// each formatted "[<constraint name>] <msg>".
void validation_messages(const Compiled& c, const std::vector<Violation>& vio, uint32_t object, std::vector<std::string>& deny,
                         std::vector<std::string>& warn) {
  auto supported = [](const std::string& a) { return a == "deny" || a == "dryrun" || a == "warn"; };
  // results are in object order: binary-search the object's range
  auto lo = std::lower_bound(vio.begin(), vio.end(), object, [](const Violation& v, uint32_t o) { return v.object < o; });
  for (auto it = lo; it != vio.end() && it->object == object; ++it) {
    const Violation& x = *it;
    const Constraint& con = *c.order[x.constraint];
    std::vector<std::string> actions;
    if (x.action == "scoped") {
      // each action listed for this enforcement point is validated on its own; unsupported ones are skipped
      VP arr = json_parse(x.scoped_json.data(), x.scoped_json.size());
      for (auto& a : arr->items)
        if (supported(a->s)) actions.push_back(a->s);
      if (actions.empty()) continue;
    } else {
      if (!supported(x.action)) continue;
      actions.push_back(x.action);
    }
    for (auto& a : actions) {
      if (a == "deny") deny.push_back("[" + con.name + "] " + x.msg);
      if (a == "warn") warn.push_back("[" + con.name + "] " + x.msg);
    }
  }
}

}  // namespace gk
