// common.hpp — shared bits of the host mirror (libnbh).
#pragma once
#include <string>

#include "nbh.h"

namespace nbh {
int fail(int code, const std::string &msg);

// Propagate a libnbk error code with its text (an OCaml stub would raise Failure here).
inline int ck(nbk_ctx *ctx, int rc) { return rc == 0 ? 0 : fail(rc, nbk_last_error(ctx)); }

#define NBH_TRY(expr)          \
  do {                         \
    int rc_ = (expr);          \
    if (rc_ != 0) return rc_;  \
  } while (0)

// base.ml:36-58
inline double norm_squared(const double *v) {
  double x = v[0], y = v[1], z = v[2];
  return x * x + y * y + z * z;
}
double norm(const double *v);
double distance(const double *x, const double *y);
inline double square(double x) { return x * x; }
}  // namespace nbh
