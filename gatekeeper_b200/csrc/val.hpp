// Value model shared by the Rego front-end, the flattener and the result materialiser.
//
// JSON documents (unstructured.Unstructured objects, constraint specs) and Rego values (which add sets and
// non-string object keys) use one immutable, ref-counted node type.  Ordering/equality follow OPA's total
// order (null < bool < number < string < array < object < set); `%v` rendering follows ast.Value.String().
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace gk {

// Device-visible value-type codes (one byte per cell in a VT column). Keep in sync with vm_core.h.
enum class VT : uint8_t { Undef = 0, Null = 1, False = 2, True = 3, Num = 4, Str = 5, Arr = 6, Obj = 7, Set = 8 };

struct Num {
  bool is_int = true;
  __int128 i = 0;
  double d = 0.0;
  static Num of_int(__int128 v) { Num n; n.is_int = true; n.i = v; n.d = (double)v; return n; }
  static Num of_double(double v);
  double as_double() const { return is_int ? (double)i : d; }
};
int num_cmp(const Num& a, const Num& b);
std::string num_str(const Num& n);
bool num_fits_i64(const Num& n, int64_t* out);
// Order-preserving int64 key of a number for the device's numeric columns and constants: 2*floor(x) + (x is not an integer),
// saturated at +-(2^63 - 2).  For an INTEGER constant k with |k| <= 2^61:  x <cmp> k  <=>  num_key(x) <cmp> num_key(k)  for every
// number x (fractions and out-of-range values included), so no object number is ever refused or approximated.
// INT64_MIN / INT64_MAX stay free: they mark non-numbers below / above every number in OPA's cross-type order.
int64_t num_key(const Num& n);
#define GK_NUM_KEY_CONST_LIMIT ((int64_t)1 << 61)

struct Node;
using VP = std::shared_ptr<const Node>;

struct Node {
  VT t = VT::Null;
  Num n;
  std::string s;
  std::vector<VP> items;                    // Arr (ordered) / Set (sorted, unique)
  std::vector<std::pair<VP, VP>> kv;        // Obj (sorted by key, unique)
};

// Nodes come from a per-thread free list of fixed-size blocks (a document is a few hundred short-lived nodes; the general
// allocator's lock-free fast path is still several times the cost of a list pop).  Blocks freed on another thread simply
// join that thread's list; a thread's list is handed to a global pool when the thread ends.
std::shared_ptr<Node> new_node();

VP v_null();
VP v_bool(bool b);
VP v_num(const Num& n);
VP v_int(long long i);
VP v_str(std::string s);
VP v_arr(std::vector<VP> items);
VP v_set(std::vector<VP> items);             // sorts + dedups
VP v_obj(std::vector<std::pair<VP, VP>> kv); // sorts; later duplicates win

VP v_deep_copy(const VP& v);                 // a copy that shares no node with `v` (thread-private documents)
int v_cmp(const VP& a, const VP& b);         // total order; both non-null
inline bool v_eq(const VP& a, const VP& b) { return v_cmp(a, b) == 0; }
int type_rank(VT t);

VP obj_get(const VP& o, const VP& key);      // nullptr when absent / not an object
VP obj_get(const VP& o, const char* key);
VP set_find(const VP& s, const VP& x);       // element or nullptr
inline bool truthy(const VP& v) { return v && v->t != VT::False; }

// ---- JSON
struct JsonError { std::string msg; };
VP json_parse(const char* p, size_t n);                    // throws JsonError
std::string json_str(const VP& v);                          // canonical: sorted keys, sets as arrays
void json_quote(const std::string& s, std::string& out);

// ---- OPA `%v`
std::string fmt_value(const VP& v, bool top);

// Canonical intern key of a value: one type char + payload; equal keys <=> v_eq.
//   's'+raw string | 'n'+canonical number | 't' | 'f' | 'z'(null) | 'j'+canonical JSON-ish (composites)
std::string intern_key(const VP& v);

}  // namespace gk
