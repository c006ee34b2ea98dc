// oracle/num.hh — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/base.h header).
//
// Number types for the integrator restatements.  The reference writes Hermite twice, once
// over float (hermite.ml) and once "line for line" over Diff.DFloat.diff
// (dFloat_hermite.ml, dFloat_advancer.ml:20-86); here one template serves both:
// T = double reproduces hermite.ml's arithmetic (same association order as base.h), and
// T = Dual is a first-order forward-mode dual number standing in for Diff.DFloat.
//
// Diff.DFloat is a third-party library that is NOT in /root/reference and is pinned nowhere
// (no opam/META file; dFloat.ml:20 just does `include Diff.DFloat`).  What is restated here
// is the textbook algorithm its call sites rely on: value + one tangent, sum/product/quotient
// rules, d sqrt(x) = dx / (2 sqrt x), d x^c = c x^(c-1) dx, comparisons on the value.
// PARITY UNPINNED for the tangent path; checked against central finite
// differences and the closed forms of SURVEY.md §8 a14 in tests/.
#pragma once
#include <cmath>

struct Dual {
  double v, d;
  Dual() : v(0.0), d(0.0) {}
  Dual(double v_) : v(v_), d(0.0) {}  // the `C x` constructor: constant, zero tangent
  Dual(double v_, double d_) : v(v_), d(d_) {}
};
inline Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
inline Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
inline Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
inline Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
inline Dual operator/(Dual a, Dual b) {
  double q = a.v / b.v;
  return Dual(q, (a.d - q * b.d) / b.v);
}
inline Dual sqrt(Dual a) {
  double s = std::sqrt(a.v);
  return Dual(s, a.d / (2.0 * s));
}
inline Dual pow(Dual a, double c) {
  return Dual(std::pow(a.v, c), c * std::pow(a.v, c - 1.0) * a.d);
}
inline bool operator<(Dual a, Dual b) { return a.v < b.v; }
inline bool operator>(Dual a, Dual b) { return a.v > b.v; }
inline bool operator==(Dual a, Dual b) { return a.v == b.v; }
inline double value_of(Dual a) { return a.v; }
inline double value_of(double a) { return a; }

namespace onum {
using std::pow;
using std::sqrt;

template <class T> inline T square(T x) { return x * x; }
// base.ml:36-58 / dFloat_advancer.ml:33-55
template <class T> inline T norm_squared(const T v[3]) {
  T x = v[0], y = v[1], z = v[2];
  return x * x + y * y + z * z;
}
template <class T> inline T norm(const T v[3]) { return sqrt(norm_squared(v)); }
template <class T> inline T distance_squared(const T x[3], const T y[3]) {
  T dx = x[0] - y[0], dy = x[1] - y[1], dz = x[2] - y[2];
  return dx * dx + dy * dy + dz * dz;
}
template <class T> inline T distance(const T x[3], const T y[3]) {
  return sqrt(distance_squared(x, y));
}
template <class T> inline T dot(const T x[3], const T y[3]) {
  return x[0] * y[0] + x[1] * y[1] + x[2] * y[2];
}
// base.ml:73-80 / dFloat_advancer.ml:57-64
template <class T>
inline void grad_v(T eps2, T m1, const T q1[3], T m2, const T q2[3], T gv[3]) {
  T r2 = distance_squared(q1, q2);
  T re = r2 + eps2;
  T r3 = re * sqrt(re);
  T factor = m1 * m2 / r3;
  gv[0] = factor * (q1[0] - q2[0]);
  gv[1] = factor * (q1[1] - q2[1]);
  gv[2] = factor * (q1[2] - q2[2]);
}
// base.ml:91-110 / dFloat_advancer.ml:66-81
template <class T>
inline void grad_v_dgrad_v(T eps2, T m1, const T q1[3], const T p1[3], T m2, const T q2[3],
                           const T p2[3], T gv[3], T dgv[3]) {
  for (int i = 0; i < 3; ++i) {
    gv[i] = q1[i] - q2[i];
    dgv[i] = p1[i] / m1 - p2[i] / m2;
  }
  T r2 = norm_squared(gv);
  T rdv = dot(gv, dgv);
  T re = r2 + eps2;
  T r3 = sqrt(re) * re;
  T r5 = r3 * re;
  T factorg = m1 * m2 / r3;
  T factord = T(3.0) * rdv * m1 * m2 / r5;
  for (int i = 0; i < 3; ++i) {
    dgv[i] = factorg * dgv[i] - factord * gv[i];
    gv[i] = factorg * gv[i];
  }
}
}  // namespace onum
