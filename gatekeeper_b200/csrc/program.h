// Device-visible layout of (a) a flattened object batch and (b) the lowered constraint table.
// Plain C structs shared by the host builder (C++), the CUDA kernels and the test-only host emulation.
//
//   batch  : column-wise (SoA) -- header columns one row per object, CSR label table, CSR "scopes" (one
//            row per iterated element, e.g. spec.containers[_]) with fixed-width feature columns.
//   program: DISTINCT match blocks (the spec.match pre-filter, pkg/mutation/match/match.go:32-65), a
//            constraint table (match block id + entry pc) ordered so that constraints sharing a match block
//            are adjacent, and one shared instruction array of warp-uniform postfix predicate code.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define GK_HD __host__ __device__ __forceinline__
#else
#define GK_HD inline
#endif

// ---- value-type codes stored in VT columns (== gk::VT)
enum { GK_VT_UNDEF = 0, GK_VT_NULL = 1, GK_VT_FALSE = 2, GK_VT_TRUE = 3, GK_VT_NUM = 4, GK_VT_STR = 5,
       GK_VT_ARR = 6, GK_VT_OBJ = 7, GK_VT_SET = 8, GK_VT_NUM_INEXACT = 9 /* number not representable as i64 */ };

// ---- column encodings (bitmask)
enum { GK_ENC_VT = 1, GK_ENC_SID = 2, GK_ENC_NUM = 4, GK_ENC_BYTES = 8 };

#define GK_SID_UNDEF 0u          /* intern id 0 is reserved: "no value" */
#define GK_NONE 0xFFFFFFFFu
#define GK_MAX_LOOP_DEPTH 4
#define GK_MAX_STACK 60          /* boolean stack lives in one 64-bit register per thread */
#define GK_MAX_CSE 64            /* shared sub-formula results live in one 64-bit register per thread */
#define GK_PC_ACCEPT 0xFFFFFFFFu
#define GK_PC_REJECT 0xFFFFFFFEu

typedef struct {
  int32_t scope;           // 0 = root (one row per object)
  uint32_t enc;
  const uint8_t* vt;       // [rows]
  const uint32_t* sid;     // [rows]
  const int64_t* num;      // [rows]
  const uint32_t* boff;    // [rows+1]
  const uint8_t* bytes;
} GkColumn;

typedef struct {
  int32_t parent;          // parent scope id (0 = root)
  uint32_t rows;
  const uint32_t* off;     // [parent_rows+1]
} GkScope;

// header flags (per object)
enum { GK_F_HAS_OBJ = 1, GK_F_IS_NS = 2, GK_F_HAS_NS = 4 /* metadata.namespace != "" */,
       GK_F_NS_OBJ = 8 /* a Namespace object is known for the review */, GK_F_SRC_SHIFT = 4, GK_F_SRC_MASK = 0x70,
       GK_F_SKIP = 128 /* review-level error (bad JSON, ...): the kernel skips the object */ };
enum { GK_SRC_EMPTY = 0, GK_SRC_ORIGINAL = 1, GK_SRC_GENERATED = 2, GK_SRC_ALL = 3, GK_SRC_INVALID = 4 };

typedef struct {
  uint32_t n;                    // objects in the batch
  uint32_t has_old;              // 1: header arrays carry 2n rows, rows [n,2n) describe OldObject
  const uint32_t* flags;         // [n or 2n]
  const uint32_t* kind_sid;      // [..]
  const uint32_t* group_sid;
  const uint32_t* nsname_sid;    // namespace NAME used by namespaces/excludedNamespaces (match.go:118-179); GK_NONE = none
  const uint32_t* name_off;      // [rows+1] metadata.name bytes
  const uint8_t* name_bytes;
  const uint32_t* gen_off;       // [rows+1] metadata.generateName bytes
  const uint8_t* gen_bytes;
  const uint32_t* lbl_off;       // [rows+1]
  const uint32_t* lbl_kv;        // 2 * nlabels: (key sid, value sid)
  const uint32_t* nsrow;         // [n] row in the namespace table or GK_NONE   (shared by obj and old)
  const uint32_t* nsl_off;       // namespace table: [nsrows+1]
  const uint32_t* nsl_kv;
  // interned-string dictionary (engine-global, append-only): bytes of sid i are dict_bytes[dict_off[i], dict_off[i+1])
  const uint32_t* dict_off;
  const uint8_t* dict_bytes;
  uint32_t dict_n;
  const GkColumn* cols;
  const GkScope* scopes;
  uint32_t ncols, nscopes;
} GkBatch;

// ---- match block
enum { GK_M_HAS_MATCH = 1, GK_M_SCOPE_CLUSTER = 2, GK_M_SCOPE_NAMESPACED = 4, GK_M_HAS_LSEL = 8, GK_M_HAS_NSSEL = 16,
       GK_M_LSEL_INVALID = 32, GK_M_NSSEL_INVALID = 64, GK_M_SRC_INVALID = 128, GK_M_HAS_NAME = 256,
       GK_M_SRC_SHIFT = 12 /* 3 bits: GK_SRC_* of the matcher (EMPTY => All) */ };
enum { GK_W_EXACT = 0, GK_W_PREFIX = 1, GK_W_SUFFIX = 2, GK_W_CONTAINS = 3 };
enum { GK_SEL_IN = 0, GK_SEL_NOTIN = 1, GK_SEL_EXISTS = 2, GK_SEL_NOTEXISTS = 3 };

// error codes written to the error list (host renders the reference's error text from them)
enum { GK_E_NONE = 0, GK_E_LSEL_INVALID = 1, GK_E_NSSEL_INVALID = 2, GK_E_NS_MISSING = 3, GK_E_SRC_INVALID_MATCH = 4,
       GK_E_SRC_UNSPECIFIED = 5, GK_E_SRC_INVALID_OBJ = 6, GK_E_NO_OBJECT = 7, GK_E_NUM_RANGE = 8 };

typedef struct {
  uint32_t flags;
  uint32_t kinds_off, kinds_n;    // pool: per entry [nk, ng, wild(bit0 kind '*', bit1 group '*'), nk kind sids, ng group sids]
  uint32_t ns_off, ns_n;          // pool: per pattern [mode, byte_off, len]
  uint32_t exns_off, exns_n;
  uint32_t lsel_off, lsel_n;      // pool: per requirement [key sid, op, nvals, vals...]
  uint32_t nssel_off, nssel_n;
  uint32_t name_mode, name_boff, name_len;
  uint32_t pad0, pad1;
} GkMatch;

typedef struct {
  uint32_t match_id;              // index into the distinct match blocks
  uint32_t pc;                    // entry pc, or GK_PC_ACCEPT / GK_PC_REJECT for constant predicates
} GkCons;

// ---- predicate instructions: 4 x u32, executed by ALL lanes of a warp in lock step (the pc is warp-uniform).
// Every lane owns a boolean stack held in one 64-bit register (bit 0 = top) and a 64-bit register of shared
// sub-formula results.  There are no data-dependent jumps: loops run for the warp-wide maximum trip count with
// finished lanes masked off, so the only divergence left is inside byte-string comparisons.
//   w0 = op | slot<<8 | col<<16      (slot: which open loop supplies the row; 0 = the object itself)
//   w1 = operand A (immediate / pool offset / scope id / cse bit)
//   w2 = jump target (loops)
//   w3 = operand B (count / length / compare op)
enum {
  GK_OP_END = 0,        // result = top of stack
  GK_OP_TRUTHY = 1,     // push vt != undef && vt != false
  GK_OP_DEFINED = 2,    // push vt != undef
  GK_OP_VTMASK = 3,     // push (1 << vt) & w1
  GK_OP_SID_EQ = 4,     // push sid == w1
  GK_OP_SID_IN = 5,     // push sid in pool[w1 .. w1+w3) (sorted)
  GK_OP_NUM_CMP = 6,    // w3 = GK_CMP_*; i64 constant at pool[w1], pool[w1+1] (lo, hi); OPA cross-type ordering
  GK_OP_PREFIX = 7,     // push vt == str && bytes startswith cbytes[w1 .. w1+w3)
  GK_OP_SUFFIX = 8,
  GK_OP_CONTAINS = 9,
  GK_OP_ANYPREFIX = 10, // pool[w1 ..]: w3 entries of [byte_off, len]
  GK_OP_ANYSUFFIX = 11,
  GK_OP_LOOP_BEGIN = 12, // w1 = scope id; slot = new loop slot; col field = parent slot; pushes acc = false;
                         // trip = warp max of the lane ranges; if trip == 0 jump to w2 (just past LOOP_END)
  GK_OP_LOOP_END = 13,   // acc |= top & lane-still-in-range; pop; ++iter; if --trip jump to w2 (loop body)
  GK_OP_AND = 14,
  GK_OP_OR = 15,
  GK_OP_NOT = 16,
  GK_OP_PUSH = 17,       // push (w1 & 1)
  GK_OP_CSE_TRY = 18,    // if shared result w1 is already valid for this warp: push it and jump to w2
  GK_OP_CSE_STORE = 19,  // shared result w1 = top (stays on the stack); mark valid
};
enum { GK_CMP_LT = 0, GK_CMP_LE = 1, GK_CMP_GT = 2, GK_CMP_GE = 3, GK_CMP_EQ = 4, GK_CMP_NE = 5 };

typedef struct { uint32_t w0, w1, w2, w3; } GkInstr;

typedef struct {
  uint32_t nconstraints;
  uint32_t nmatch;
  uint32_t ninstr;
  uint32_t npool;
  uint32_t ncbytes;
  const GkCons* cons;        // [nconstraints]
  const GkMatch* match;      // [nmatch]
  const GkInstr* instr;      // [ninstr]
  const uint32_t* pool;      // u32 constant pool
  const uint8_t* cbytes;     // constant byte strings (wildcard literals, prefixes)
} GkProgram;

typedef struct {
  uint32_t* viol;            // [n * words]   bit c%32 of word c/32 set: object violates constraint c
  uint32_t* err;             // [n * words]   matcher error ("autoreject") plane
  unsigned long long* totals;     // [nconstraints] violating pairs
  unsigned long long* err_totals; // [nconstraints]
  uint32_t* errlist;         // [errcap * 3]  (object, constraint, code)
  uint32_t* errcount;        // [1]
  uint32_t errcap;
  uint32_t words;            // ceil(nconstraints / 32)
} GkOut;
