/* C mirror of go/gpudriver/driver.go (*Driver).reviewBatch: the same C-ABI calls in the same order with the same field
 * population (a C-allocated gk_obj array; json / old_json / ns_json / ns_name / operation / userinfo_json / source set exactly
 * when the Go shim sets them; object_errors read before the results), since this image cannot compile the Go file itself.
 *
 *   mirror <lib.so> <input file>      prints one line per result:  <object>\t<Kind/name>\t<action>\t<autoreject>\t<msg>
 *                                     or "ERR\t<object>\t<text>" for a review-level error
 * Input (written by tests/test_c_abi.py), all integers decimal, one item per line, payloads length-prefixed:
 *   T <kind> <len>\n<rego bytes>\n     C <len>\n<constraint json>\n     N <name> <len>\n<namespace json>\n     E <enforcement point>\n
 *   R <source> <operation or -> <ns_name or -> <len obj> <len old> <len ns> <len userinfo>\n<obj><old><ns><userinfo>\n
 *   X <len>\n<ExpansionTemplate json>\n   (UpsertExpansionTemplate: a refusal is printed as "XERR\t<text>", not fatal -- the shim returns it)
 * With an X line in the input the run ends with "CONFLICTS\t<json array>" (ExpansionConflicts).
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/gk_engine.h"

#define SYM(name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

static char* read_n(FILE* f, size_t n) {
  char* b = (char*)malloc(n + 1);
  if (n && fread(b, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  b[n] = 0;
  fgetc(f); /* newline */
  return b;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  SYM(gk_engine_create) SYM(gk_engine_destroy) SYM(gk_add_template) SYM(gk_add_constraint) SYM(gk_put_namespace) SYM(gk_add_data) SYM(gk_review_batch)
  SYM(gk_result_constraint_key) SYM(gk_free_result) SYM(gk_free_str) SYM(gk_add_expansion_template) SYM(gk_expansion_conflicts)
  int any_x = 0;
  FILE* f = fopen(argv[2], "rb");
  if (!f) return 2;
  char* err = NULL;
  gk_cfg cfg;
  memset(&cfg, 0, sizeof cfg);
  gk_engine_t* e = p_gk_engine_create(&cfg, &err);
  if (!e) { fprintf(stderr, "create: %s\n", err ? err : "?"); return 3; }
  char ep[128] = "";
  size_t cap = 64, n = 0;
  gk_obj* objs = (gk_obj*)calloc(cap, sizeof(gk_obj));   /* C memory, like the shim's C.calloc */
  char tag;
  while (fscanf(f, " %c", &tag) == 1) {
    if (tag == 'T') {
      char kind[256]; size_t len;
      if (fscanf(f, "%255s %zu", kind, &len) != 2) return 2;
      fgetc(f);
      char* rego = read_n(f, len);
      if (p_gk_add_template(e, kind, rego, len, &err)) { fprintf(stderr, "template: %s\n", err); return 3; }
      free(rego);
    } else if (tag == 'C') {
      size_t len;
      if (fscanf(f, "%zu", &len) != 1) return 2;
      fgetc(f);
      char* js = read_n(f, len);
      if (p_gk_add_constraint(e, js, len, &err)) { fprintf(stderr, "constraint: %s\n", err); return 3; }
      free(js);
    } else if (tag == 'N') {
      char name[256]; size_t len;
      if (fscanf(f, "%255s %zu", name, &len) != 2) return 2;
      fgetc(f);
      char* js = read_n(f, len);
      if (p_gk_put_namespace(e, name, js, len, &err)) { fprintf(stderr, "namespace: %s\n", err); return 3; }
      {   /* (driver.go AddData: the object also goes to data.inventory under its ProcessData path) */
        const char* path[4] = {"cluster", "v1", "Namespace", name};
        if (p_gk_add_data(e, path, 4, js, len, &err)) { fprintf(stderr, "data: %s\n", err); return 3; }
      }
      free(js);
    } else if (tag == 'X') {
      size_t len;
      if (fscanf(f, "%zu", &len) != 1) return 2;
      fgetc(f);
      char* js = read_n(f, len);
      any_x = 1;
      if (p_gk_add_expansion_template(e, js, len, &err)) {
        printf("XERR\t%s\n", err ? err : "?");
        p_gk_free_str(err);
        err = NULL;
      }
      free(js);
    } else if (tag == 'E') {
      if (fscanf(f, "%127s", ep) != 1) return 2;
    } else if (tag == 'R') {
      int source; char op[64], nsn[256]; size_t lo, ll, ln, lu;
      if (fscanf(f, "%d %63s %255s %zu %zu %zu %zu", &source, op, nsn, &lo, &ll, &ln, &lu) != 7) return 2;
      fgetc(f);
      char* all = read_n(f, lo + ll + ln + lu);
      if (n == cap) { cap *= 2; objs = (gk_obj*)realloc(objs, cap * sizeof(gk_obj)); memset(objs + n, 0, (cap - n) * sizeof(gk_obj)); }
      gk_obj* o = &objs[n++];
      /* the population rules of reviewBatch: a field stays NULL when the AdmissionRequest does not carry it */
      if (lo) { o->json = all; o->len = lo; }
      if (ll) { o->old_json = all + lo; o->old_len = ll; }
      if (ln) { o->ns_json = all + lo + ll; o->ns_len = ln; }
      if (strcmp(nsn, "-")) o->ns_name = strdup(nsn);
      if (strcmp(op, "-")) o->operation = strdup(op);
      if (lu) { o->userinfo_json = all + lo + ll + ln; o->userinfo_len = lu; }
      o->source = (uint8_t)source;
    }
  }
  gk_result res;
  if (p_gk_review_batch(e, objs, n, ep, GK_F_MATERIALIZE, &res, &err)) { fprintf(stderr, "review: %s\n", err); return 3; }
  if (res.object_errors)
    for (size_t i = 0; i < n; ++i)
      if (res.object_errors[i]) printf("ERR\t%zu\t%s\n", i, res.object_errors[i]);
  for (size_t i = 0; i < res.n_violations; ++i) {
    const gk_violation* v = &res.violations[i];
    printf("%u\t%s\t%s\t%d\t%s\n", v->object, p_gk_result_constraint_key(&res, v->constraint), v->enforcement_action, (int)v->autoreject, v->msg);
  }
  p_gk_free_result(&res);
  if (any_x) {
    char* cj = p_gk_expansion_conflicts(e);
    printf("CONFLICTS\t%s\n", cj ? cj : "[]");
    p_gk_free_str(cj);
  }
  p_gk_engine_destroy(e);
  return 0;
}
