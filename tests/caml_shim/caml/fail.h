#ifndef SHIM_CAML_FAIL_H
#define SHIM_CAML_FAIL_H
#include "mlvalues.h"
/* raise Failure msg: never returns (the shim unwinds to shim_invoke with longjmp) */
void caml_failwith(const char *msg) __attribute__((noreturn));
void caml_invalid_argument(const char *msg) __attribute__((noreturn));
#endif
