/* tests/caml_shim/shim.c - TEST INFRASTRUCTURE: the few runtime functions ocaml/nbk_stubs.c calls,
 * plus what a test needs to build OCaml values and to call a stub the way ocamlopt (all arguments
 * in registers) and the bytecode interpreter (argv, argn - only for more than 5 arguments) do.
 * See caml/mlvalues.h. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/memory.h>

static jmp_buf shim_env;
static int shim_env_set = 0;
static char shim_error[1024];

value caml_alloc(mlsize_t wosize, tag_t tag) {
  header_t *b = (header_t *)calloc(wosize + 1, sizeof(value));
  b[0] = ((header_t)wosize << 10) | tag;
  value v = (value)(b + 1);
  if (tag < 251)
    for (mlsize_t i = 0; i < wosize; ++i) Field(v, i) = Val_unit;
  return v;
}
value caml_alloc_tuple(mlsize_t n) { return caml_alloc(n, 0); }
value caml_copy_double(double d) {
  value v = caml_alloc(1, Double_tag);
  Double_val(v) = d;
  return v;
}
value caml_copy_string(const char *s) {
  size_t len = strlen(s), words = len / sizeof(value) + 1;
  value v = caml_alloc(words, 252);
  memcpy((char *)v, s, len);
  ((char *)v)[words * sizeof(value) - 1] = (char)(words * sizeof(value) - 1 - len);
  return v;
}
value caml_alloc_custom(struct custom_operations *ops, uintnat size, mlsize_t mem, mlsize_t max) {
  (void)mem; (void)max;
  value v = caml_alloc(1 + (size + sizeof(value) - 1) / sizeof(value), Custom_tag);
  Custom_ops_val(v) = ops;
  return v;
}
static struct custom_operations shim_ba_ops = {"_bigarr02", NULL, NULL, NULL, NULL, NULL, NULL, NULL};
value caml_ba_alloc_dims(int flags, int num_dims, void *data, ...) {
  value v = caml_alloc_custom(&shim_ba_ops, sizeof(struct caml_ba_array) + 4 * sizeof(intnat), 0, 1);
  struct caml_ba_array *b = Caml_ba_array_val(v);
  va_list ap;
  va_start(ap, data);
  b->data = data;
  b->num_dims = num_dims;
  b->flags = flags;
  b->proxy = NULL;
  for (int i = 0; i < num_dims; ++i) b->dim[i] = va_arg(ap, intnat);
  va_end(ap);
  return v;
}
void caml_failwith(const char *msg) {
  snprintf(shim_error, sizeof shim_error, "Failure: %s", msg ? msg : "(null)");
  if (!shim_env_set) { fprintf(stderr, "uncaught %s\n", shim_error); abort(); }
  longjmp(shim_env, 1);
}
void caml_invalid_argument(const char *msg) {
  snprintf(shim_error, sizeof shim_error, "Invalid_argument: %s", msg ? msg : "(null)");
  if (!shim_env_set) { fprintf(stderr, "uncaught %s\n", shim_error); abort(); }
  longjmp(shim_env, 1);
}

/* ---- helpers for the Python side --------------------------------------------------------------- */
value shim_val_int(long n) { return Val_long(n); }
long shim_int_val(value v) { return Long_val(v); }
value shim_val_double(double d) { return caml_copy_double(d); }
double shim_double_val(value v) { return Double_val(v); }
value shim_tuple(int n, const value *fields) {
  value v = caml_alloc_tuple((mlsize_t)n);
  for (int i = 0; i < n; ++i) Field(v, i) = fields[i];
  return v;
}
value shim_field(value v, int i) { return Field(v, i); }
int shim_is_block(value v) { return Is_block(v); }
long shim_wosize(value v) { return (long)Wosize_val(v); }
int shim_tag(value v) { return (int)Tag_val(v); }
/* a Bigarray.Array1 over the caller's memory: kind 1 = float64, 6 = int32 */
value shim_bigarray1(int kind, void *data, long dim) {
  return caml_ba_alloc_dims(kind | CAML_BA_C_LAYOUT, 1, data, (intnat)dim);
}
void shim_finalize(value v) {
  if (Is_block(v) && Tag_val(v) == Custom_tag && Custom_ops_val(v)->finalize) Custom_ops_val(v)->finalize(v);
}
const char *shim_last_error(void) { return shim_error; }

typedef value (*fn1)(value);
typedef value (*fn2)(value, value);
typedef value (*fn3)(value, value, value);
typedef value (*fn4)(value, value, value, value);
typedef value (*fn5)(value, value, value, value, value);
typedef value (*fn6)(value, value, value, value, value, value);
typedef value (*fn7)(value, value, value, value, value, value, value);
typedef value (*fn8)(value, value, value, value, value, value, value, value);
typedef value (*fn9)(value, value, value, value, value, value, value, value, value);
typedef value (*fn10)(value, value, value, value, value, value, value, value, value, value);
typedef value (*fn11)(value, value, value, value, value, value, value, value, value, value, value);
typedef value (*fn12)(value, value, value, value, value, value, value, value, value, value, value, value);
typedef value (*fnbc)(value *, int);

/* Calls a stub with n arguments: bytecode = 0 the native convention, 1 the (argv, argn) one.
 * Returns 0 and *result, or 1 after an OCaml exception (message in shim_last_error). */
int shim_invoke(void *fn, int bytecode, int n, value *a, value *result) {
  shim_error[0] = 0;
  if (setjmp(shim_env)) { shim_env_set = 0; return 1; }
  shim_env_set = 1;
  value r;
  if (bytecode) r = ((fnbc)fn)(a, n);
  else switch (n) {
    case 1: r = ((fn1)fn)(a[0]); break;
    case 2: r = ((fn2)fn)(a[0], a[1]); break;
    case 3: r = ((fn3)fn)(a[0], a[1], a[2]); break;
    case 4: r = ((fn4)fn)(a[0], a[1], a[2], a[3]); break;
    case 5: r = ((fn5)fn)(a[0], a[1], a[2], a[3], a[4]); break;
    case 6: r = ((fn6)fn)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
    case 7: r = ((fn7)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
    case 8: r = ((fn8)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
    case 9: r = ((fn9)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
    case 10: r = ((fn10)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]); break;
    case 11: r = ((fn11)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10]); break;
    case 12: r = ((fn12)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11]); break;
    default: shim_env_set = 0; snprintf(shim_error, sizeof shim_error, "shim_invoke: %d arguments", n); return 1;
  }
  shim_env_set = 0;
  *result = r;
  return 0;
}
