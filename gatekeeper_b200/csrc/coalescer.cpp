// Admission coalescer: concurrent single-review callers (the webhook's request goroutines, one cgo call each) are gathered
// into micro-batches and evaluated with ONE gk_review_batch -- the reference reviews each request on its own
// (pkg/webhook/policy.go:580-675, Review at :661); a GPU wants a batch.
//
// Leader / follower, no background thread: the caller that finds the queue empty becomes the leader of the next batch, waits
// until `max_batch` reviews have joined or `max_wait_us` have passed, takes the batch, runs it, and hands every follower its
// own request's outcome (deny / warn message lists as validationHandler.getValidationMessages builds them, plus the raw
// results).  Built on the public C ABI only (include/gk_engine.h).
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gk_engine.h"
#include "val.hpp"

namespace {

struct Ticket {
  gk_obj obj;
  std::string out_json, error;
  bool done = false;
};

struct Batch {
  std::vector<Ticket*> tickets;
  bool closed = false;   // the leader has taken it: no more joiners
};

}  // namespace

struct gk_coalescer {
  gk_engine_t* e = nullptr;
  uint32_t max_batch = 64, max_wait_us = 200, flags = 0;
  std::string ep;
  std::mutex mu;
  std::condition_variable cv_full;   // wakes the leader when the batch fills
  std::condition_variable cv_done;   // wakes followers when their batch has been evaluated
  std::shared_ptr<Batch> open;       // the batch currently accepting joiners (null: the next caller leads)
  uint64_t batches = 0, reviews = 0;
};

static char* dup_c(const std::string& s) {
  char* p = (char*)malloc(s.size() + 1);
  if (p) memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

static void run_batch(gk_coalescer* c, Batch& b) {
  const size_t n = b.tickets.size();
  std::vector<gk_obj> objs(n);
  for (size_t i = 0; i < n; ++i) objs[i] = b.tickets[i]->obj;
  gk_result res;
  char* err = nullptr;
  int rc = gk_review_batch(c->e, objs.data(), n, c->ep.c_str(), c->flags | GK_F_MATERIALIZE, &res, &err);
  if (rc != 0) {
    std::string msg = err ? err : "review failed";
    gk_free_str(err);
    for (auto* t : b.tickets) t->error = msg;
    return;
  }
  // results are in object order (gk_review_batch sorts them so, also after folding expansion results onto their parents); a
  // result that is not would be lost below, so check rather than assume
  for (size_t k = 1; k < res.n_violations; ++k)
    if (res.violations[k].object < res.violations[k - 1].object) {
      for (auto* t : b.tickets) t->error = "internal: review results are not in object order";
      gk_free_result(&res);
      return;
    }
  size_t v = 0;
  for (size_t i = 0; i < n; ++i) {
    Ticket& t = *b.tickets[i];
    std::string o = "{\"batch_size\":" + std::to_string(n) + ",\"error\":";
    const char* oe = res.object_errors ? res.object_errors[i] : nullptr;
    if (oe) gk::json_quote(oe, o);
    else o += "null";
    char* vm_err = nullptr;
    char* vm = gk_validation_messages(c->e, &res, (uint32_t)i, &vm_err);
    o += ",\"messages\":";
    o += vm ? vm : "{\"deny\":[],\"warn\":[]}";
    gk_free_str(vm);
    gk_free_str(vm_err);
    o += ",\"results\":[";
    bool first = true;
    for (; v < res.n_violations && res.violations[v].object == i; ++v) {
      const gk_violation& x = res.violations[v];
      if (!first) o += ",";
      first = false;
      o += "{\"constraint\":";
      {
        const char* key = gk_result_constraint_key(&res, x.constraint);   // (the result's own snapshot: constraints may change meanwhile)
        gk::json_quote(key ? key : "", o);
      }
      o += ",\"msg\":";
      gk::json_quote(x.msg, o);
      o += ",\"details\":";
      o += (x.details_json && x.details_json[0]) ? x.details_json : "null";
      o += ",\"enforcementAction\":";
      gk::json_quote(x.enforcement_action, o);
      o += ",\"scopedEnforcementActions\":";
      o += (x.scoped_actions_json && x.scoped_actions_json[0]) ? x.scoped_actions_json : "[]";
      o += ",\"autoreject\":";
      o += x.autoreject ? "true" : "false";
      o += "}";
    }
    o += "]}";
    t.out_json = std::move(o);
  }
  gk_free_result(&res);
}

extern "C" {

gk_coalescer_t* gk_coalescer_create(gk_engine_t* e, uint32_t max_batch, uint32_t max_wait_us, const char* enforcement_point, uint32_t flags,
                                    char** err) {
  if (!e) {
    if (err) *err = dup_c("no engine");
    return nullptr;
  }
  auto* c = new gk_coalescer();
  c->e = e;
  c->max_batch = max_batch ? max_batch : 64;
  c->max_wait_us = max_wait_us;
  c->ep = enforcement_point ? enforcement_point : "validation.gatekeeper.sh";
  c->flags = flags;
  return c;
}

int gk_coalescer_review(gk_coalescer_t* c, const gk_obj* obj, char** out_json, char** err) {
  if (!c || !obj || !out_json) return GK_ERR_INVALID;
  Ticket t;
  t.obj = *obj;
  std::shared_ptr<Batch> mine;
  bool leader = false;
  {
    std::unique_lock<std::mutex> l(c->mu);
    if (!c->open) {
      c->open = std::make_shared<Batch>();
      leader = true;
    }
    mine = c->open;
    mine->tickets.push_back(&t);
    if (leader) {
      // wait for joiners: until the batch is full or the window closes
      auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(c->max_wait_us);
      while (mine->tickets.size() < c->max_batch && c->cv_full.wait_until(l, deadline) != std::cv_status::timeout) {
      }
      mine->closed = true;
      if (c->open == mine) c->open.reset();   // the next caller starts a new batch (it may run concurrently with this one)
      c->batches++;
      c->reviews += mine->tickets.size();
    } else {
      if (mine->tickets.size() >= c->max_batch) {
        c->open.reset();             // full: later callers must not join it
        c->cv_full.notify_all();
      }
      c->cv_done.wait(l, [&] { return t.done; });
    }
  }
  if (leader) {
    // outside the lock: other batches can form and run meanwhile.  Whatever happens inside, every follower of this batch is
    // released: an exception becomes the error of every ticket instead of unwinding past the waiters (and the C boundary).
    try {
      run_batch(c, *mine);
    } catch (std::exception& x) {
      for (auto* tk : mine->tickets)
        if (tk->out_json.empty() && tk->error.empty()) tk->error = std::string("admission micro-batch failed: ") + x.what();
    } catch (...) {
      for (auto* tk : mine->tickets)
        if (tk->out_json.empty() && tk->error.empty()) tk->error = "admission micro-batch failed";
    }
    std::lock_guard<std::mutex> l(c->mu);
    for (auto* x : mine->tickets) x->done = true;
    c->cv_done.notify_all();
  }
  if (!t.error.empty()) {
    if (err) *err = dup_c(t.error);
    return GK_ERR_BACKEND;
  }
  *out_json = dup_c(t.out_json);
  return GK_OK;
}

void gk_coalescer_stats(gk_coalescer_t* c, uint64_t* batches, uint64_t* reviews) {
  if (!c) return;
  std::lock_guard<std::mutex> l(c->mu);
  if (batches) *batches = c->batches;
  if (reviews) *reviews = c->reviews;
}

void gk_coalescer_destroy(gk_coalescer_t* c) { delete c; }

}  // extern "C"
