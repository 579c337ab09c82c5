// Ahead-of-time lowering: (template Rego, constraint parameters) -> flat predicate program + column schema.
//
// At AddConstraint time `input.parameters` is a constant, so the template's `violation` rule is partially
// evaluated with parameters concrete and `input.review` symbolic:
//   * sub-terms that depend only on parameters are folded by the concrete evaluator;
//   * sub-terms that depend only on the object become *closures*: parameter-independent feature columns
//     that the flattener extracts once per object (or per iterated element) and lays out column-wise;
//   * iteration over object collections (`spec.containers[_]`) becomes an EXISTS over a CSR scope;
//   * every comparison / membership / prefix test that mixes a column with a parameter constant becomes a
//     device atom; boolean structure (rule bodies = AND, multiple definitions = OR, `not`) becomes the
//     formula tree that is compiled to jump-threaded instructions (program.h).
// Anything outside this scheme is rejected with a rego_unsupported error -- there is no fallback path.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "program.h"
#include "rego.hpp"

namespace gk {

struct Closure;
using CP = std::shared_ptr<const Closure>;

struct CapArg {
  enum K : uint8_t { Conc, Col } k = Conc;
  VP v;
  CP col;
};

struct Closure {
  enum Leaf : uint8_t { None, Elem, Key } leaf = None;   // Elem/Key: the element / key of scope `scope`
  std::shared_ptr<const Module> mod;
  TP term;
  std::vector<std::pair<int, CapArg>> caps;             // captured variables (vid -> value)
  int scope = 0;                                         // innermost scope the value depends on (0 = root)
  bool uses_data = false;                                // reads data.inventory (directly or through a captured column)
  std::string key;                                       // canonical text: dedupes columns across constraints
};

// How a closure's value can be computed from the raw object JSON by the device ingest kernels (ingest_core.h):
//   Path  : <base>["a"]["b"]... with literal keys, rooted at `input.review...` or at another device closure
//   Elem / Key : the element / key of the enclosing scope's current row
//   Count : count(<device closure>) of an array / object / string
//   Lut   : a closed pure term over captured device closures (a builtin or pure user function of leaf values,
//           `equal(a, b)`, re_match(pattern, s) ...): the device hashes the raw bytes of the arguments and looks the value
//           up in a table the host fills once per DISTINCT argument tuple with the concrete evaluator
//   Host  : anything else -- only the host flattener can compute it (the snapshot then flattens on the host)
enum class XK : uint8_t { Host, Path, Elem, Key, Count, Lut };
struct XInfo {
  XK k = XK::Host;
  bool from_input = false;    // Path: rooted at `input` (keys[0] == "review")
  CP base;                    // Path (not from_input), Count: the closure the value is read from
  std::vector<VP> keys;       // Path: literal keys
  // Lut / Count: the LEAF values the closure reads, normalised to `<leaf>["a"]["b"]` with leaf = the element / key of an
  // enclosing scope, or (leaf == null) the `input` document.  They are the parts of the hash key; the host evaluates the
  // closure against skeleton documents that hold just these paths.
  struct Arg {
    CP leaf;
    std::vector<VP> keys;
  };
  std::vector<Arg> args;
};
XInfo closure_xinfo(const Closure& c);
// the leaf values behind <base>[keys...] (false: something on the way needs the host)
bool xinfo_leaf_args(const CP& base, const std::vector<VP>& keys, std::vector<XInfo::Arg>& out);

struct ScopeDef {
  int parent = 0;
  CP gen;          // closure producing the iterated collection (evaluated per parent row); null for root
  int depth = 0;
};

struct ColDef {
  CP expr;
  int scope = 0;
  uint32_t enc = 0;
};

// The column/scope schema shared by every constraint of an engine (rebuilt on each compile).
struct Schema {
  std::vector<ScopeDef> scopes;   // [0] = root
  std::vector<ColDef> cols;
  std::map<std::string, int> scope_ix, col_ix;
  bool uses_data = false;         // some closure reads `data` (data.inventory)
  bool device_only = false;       // lowering for the device ingest path: a scope / column the ingest kernels cannot compute is an error
  Schema() { scopes.emplace_back(); }
  int scope_for(const CP& gen);
  int col_for(const CP& expr, uint32_t enc);
};

// ---- formula tree
struct Formula;
using FP = std::shared_ptr<const Formula>;
struct Formula {
  enum K : uint8_t { True, False, And, Or, Not, Exists, Atom } k = True;
  std::vector<FP> kids;
  int scope = 0;          // Exists
  bool two = false;       // Exists: at least TWO rows satisfy the body (ambiguity formulas only)
  // Atom:
  int op = 0;             // GK_OP_*
  int col = -1;           // schema column
  VP cval;                // constant operand (string / number / set / list)
  uint32_t imm = 0;       // VTMASK mask / compare op
};
FP f_true();
FP f_false();
FP f_and(FP a, FP b);
FP f_or(FP a, FP b);
FP f_not(FP a);
FP f_exists(int scope, FP body);
FP f_exists2(int scope, FP body);
std::string formula_str(const FP& f, const Schema& s);
void check_netlist_shape(const FP& formula, const Schema& s);   // throws rego_unsupported for shapes the netlist cannot hold
size_t formula_size(const FP& f);

// Interns strings/values for SID columns and constants (engine-global, append-only).
struct Interner {
  virtual ~Interner() {}
  virtual uint32_t intern(const std::string& key) = 0;
};

// Lower one constraint's violation predicate.  Throws RegoError on unsupported constructs.
// `device_mode`: object-only sub-terms that the ingest kernels cannot compute (helper rules, comprehensions, impure
// functions) are inlined into the formula instead of becoming host closures; throws when that is not possible.
// `single_result`: set when a (constraint, object) pair can have at most one result (one path to the head, a head of parameters
// and object-level values only) -- the audit then counts the pair without evaluating it.
// `amb`: a predicate that is FALSE only where the pair has at most one result -- no two rule paths reach a head together and no
// iteration the head sits in has two satisfying rows (conservative: true means "count it on the host").
FP lower_violation(const std::shared_ptr<const Module>& mod, const VP& parameters, Schema& schema, bool device_mode = false, bool* single_result = nullptr,
                   FP* amb = nullptr);
bool schema_device_ingestable(const Schema& s, std::string* why = nullptr);

// Netlist assembly: every constraint's formula is merged into one DAG of bit-column ops (program.h GkOp).
struct NetBuilder {
  std::vector<GkOp> ops;
  std::vector<uint32_t> items;           // work items (op | part<<20 | nparts<<26), grouped by phase
  std::vector<uint32_t> phase_off;       // [nphases + 1]
  std::vector<GkOutEnt> outs;            // per constraint
  std::vector<uint8_t> slot_level;       // scope id of each shared-memory slot
  std::vector<uint32_t> pool;
  std::vector<uint8_t> cbytes;
  Interner* interner = nullptr;
  const Schema* schema = nullptr;
  size_t n_nodes = 0, n_atoms = 0, n_gates = 0, n_phases = 0;
  uint32_t add_bytes(const std::string& s);
  // formulas[c] / match_id[c] per constraint (in final constraint order); nmatch distinct match blocks
  void build(const std::vector<FP>& formulas, const std::vector<uint32_t>& match_id, uint32_t nmatch);
};

}  // namespace gk
