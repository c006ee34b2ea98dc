/* tests/caml_shim/caml/mlvalues.h - TEST INFRASTRUCTURE: a stand-in for the OCaml runtime headers.
 *
 * The build image has no OCaml, so ocaml/nbk_stubs.c could never be compiled, let alone run.  This
 * directory re-creates just enough of <caml/...> - with the value representation the OCaml manual
 * documents for 64-bit targets (ch. "Interfacing C with OCaml": tagged integers, blocks with a
 * header word, boxed doubles, custom blocks whose first word is the operations table, Bigarray's
 * caml_ba_array laid out behind it) - for the stubs to compile UNCHANGED into a shared library and
 * be called from tests/test_ocaml_stubs*.py.  It is written from the manual's description, not
 * copied from the runtime; it has no GC (blocks are malloc'ed and leaked), which is all a stub
 * that never allocates across a collection needs. */
#ifndef SHIM_CAML_MLVALUES_H
#define SHIM_CAML_MLVALUES_H
#include <stddef.h>
#include <stdint.h>

typedef intptr_t intnat;
typedef uintptr_t uintnat;
typedef intnat value;
typedef uintnat header_t;
typedef uintnat mlsize_t;
typedef unsigned int tag_t;

#define Val_long(x) ((intnat)(((uintnat)(x) << 1)) + 1)
#define Long_val(x) ((x) >> 1)
#define Val_int(x) Val_long(x)
#define Int_val(x) ((int)Long_val(x))
#define Val_unit Val_int(0)
#define Val_bool(x) Val_int((x) != 0)
#define Bool_val(x) Int_val(x)
#define Val_true Val_int(1)
#define Val_false Val_int(0)
#define Is_long(x) (((x) & 1) != 0)
#define Is_block(x) (((x) & 1) == 0)

#define Hd_val(v) (((header_t *)(v))[-1])
#define Wosize_val(v) (Hd_val(v) >> 10)
#define Tag_val(v) ((tag_t)(Hd_val(v) & 0xFF))
#define Field(v, i) (((value *)(v))[i])
#define Double_tag 253
#define Double_array_tag 254
#define Custom_tag 255
#define Double_val(v) (*(double *)(v))
#define Double_field(v, i) (((double *)(v))[i])
#define Store_double_field(v, i, d) (((double *)(v))[i] = (d))
#define Data_custom_val(v) ((void *)&Field((v), 1))

#define CAMLprim
#define CAMLextern extern
#endif
