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
}

void AuditRun::add_batch(Engine& eng, const Compiled& c, const std::vector<ObjIn>& objs, const uint32_t* viol, const uint32_t* err,
                         uint32_t words, const std::vector<uint32_t>& errlist, const std::string& ep) {
  const size_t n = objs.size();
  const uint32_t C = (uint32_t)c.order.size();
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
  std::string o = "{\"objects\":" + std::to_string(objects) + ",\"results\":" + std::to_string(results) + ",\"objectErrors\":{\"count\":" +
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
