#ifndef SHIM_CAML_ALLOC_H
#define SHIM_CAML_ALLOC_H
#include "mlvalues.h"
value caml_alloc(mlsize_t wosize, tag_t tag);
value caml_alloc_tuple(mlsize_t n);
value caml_copy_double(double d);
value caml_copy_string(const char *s);
#endif
