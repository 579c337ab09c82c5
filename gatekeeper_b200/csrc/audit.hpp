// Host-side aggregation around the batch review: audit status lists and admission messages (see audit.cpp).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "backend.hpp"
#include "engine.hpp"

namespace gk {

// StatusViolation -- pkg/audit/manager.go:99-109
struct StatusViolation {
  std::string group, version, kind, ns, name, message, action;
  std::string scoped_json;   // EnforcementActions as a JSON array ("[]" / "" when none)
};

std::string truncate_string(const std::string& s, size_t size);
bool sv_less(const StatusViolation& a, const StatusViolation& b);

// LimitQueue -- pkg/audit/manager.go:161-202: keeps the `limit` smallest violations (max-heap, largest on top)
struct LimitQueue {
  size_t limit = 20;
  std::vector<StatusViolation> heap;
  void push(StatusViolation v);
  std::vector<StatusViolation> drain_descending();
};

struct AuditRun {
  size_t limit = 20;      // --constraint-violations-limit (manager.go:64)
  size_t msg_size = 256;  // msgSize (manager.go:48)
  struct PerConstraint {
    uint64_t total = 0;
    LimitQueue queue;
  };
  std::map<std::string, PerConstraint> per_constraint;   // key: "Kind/name"
  std::map<std::string, uint64_t> by_action;
  uint64_t objects = 0, results = 0;
  // objects the review refused (undecodable JSON, kind missing ...): counted and the first few kept, so that a sweep never
  // under-reports silently (the reference logs them per object: pkg/audit/manager.go:722-729)
  uint64_t object_errors = 0, seen_objects = 0;
  std::vector<std::pair<uint64_t, std::string>> first_object_errors;   // (object index within the run, text)
  void add_object_errors(const std::vector<std::string>& errs);

  void fold(const std::string& key, StatusViolation sv);
  void merge(AuditRun& other);
  // fold every result of a reviewed batch: viol/err are the kernel's bitmaps [n * words]
  // `id` (the batch's namespace / name arrays, read back from the device): with it, the pairs of constraints that have one
  // result per pair (Compiled::single_result) are COUNTED from the bitmap, and only the objects that can still enter a
  // constraint's list -- the `limit` smallest by (namespace, name), ties included -- are evaluated for their messages.
  void add_batch(Engine& eng, const Compiled& c, const std::vector<ObjIn>& objs, const uint32_t* viol, const uint32_t* err,
                 uint32_t words, const std::vector<uint32_t>& errlist, const std::string& ep, const BatchIdentity* id = nullptr,
                 const uint32_t* amb = nullptr);   // amb: the ambiguity netlist's bitmap (bit clear = at most one result for the pair)
  uint64_t rendered_pairs = 0, counted_pairs = 0;   // pairs evaluated on the host / taken from the bitmap (lazy path)
  std::string report();
};

void validation_messages(const Compiled& c, const std::vector<Violation>& vio, uint32_t object, std::vector<std::string>& deny,
                         std::vector<std::string>& warn);

}  // namespace gk
