// oracle/leapfrog.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/base.h header).
//
// Restatement of the block-time-step DKD leapfrog and of the Omelyan/Chin-Chen advancer:
//   leapfrog.ml:64-68    drift
//   leapfrog.ml:70-104   kick (oracle/base.h ob_kick)
//   leapfrog.ml:106-117  sort_sub / find_can_advance_index
//   leapfrog.ml:119-142  advance_subsystem
//   leapfrog.ml:144-165  advance / advance_const_dt
//   omelyan_advancer.ml:39-104  iter_pairwise / b_operate / a_operate / c_operate / advance
// in the reference's sequential pair order (so the kick's time-scale half sees the partially
// kicked velocities exactly as leapfrog.ml:77-93 does).
#include "oracle.h"
extern "C" {
#include "base.h"
}

#include <algorithm>
#include <cmath>
#include <vector>

namespace {

struct LBody {
  int id;
  double t, dt_max, m, q[3], v[3];
};

// OCaml's polymorphic `compare` on floats: nan equals nan and is below everything else.
int ocaml_compare(double x, double y) {
  if (x < y) return -1;
  if (x > y) return 1;
  if (x == y) return 0;
  bool xn = std::isnan(x), yn = std::isnan(y);
  if (xn && yn) return 0;
  return xn ? -1 : 1;
}

void drift(double dt, LBody &b) {
  b.t = b.t + dt;
  b.q[0] = b.q[0] + dt * b.v[0];
  b.q[1] = b.q[1] + dt * b.v[1];
  b.q[2] = b.q[2] + dt * b.v[2];
}

void kick(double eta, double dt, LBody &b1, LBody &b2) {
  ob_kick(eta, dt, b1.m, b1.q, b1.v, &b1.dt_max, b2.m, b2.q, b2.v, &b2.dt_max);
}

void sort_sub(std::vector<LBody> &bs, int iend) {
  std::stable_sort(bs.begin(), bs.begin() + iend, [](const LBody &a, const LBody &b) {
    return ocaml_compare(a.dt_max, b.dt_max) < 0;
  });
}

int find_can_advance_index(const std::vector<LBody> &bs, int iend, double dt) {
  while (!(iend == 0 || bs[iend - 1].dt_max < dt)) --iend;
  return iend;
}

void advance_subsystem(double dt, double eta, std::vector<LBody> &bs, int iend, long *nkicks) {
  if (iend <= 0) return;
  int ibegin = find_can_advance_index(bs, iend, dt);
  double dt2 = dt / 2.0;
  for (int i = ibegin; i < iend; ++i) bs[i].dt_max = INFINITY;
  advance_subsystem(dt2, eta, bs, ibegin, nkicks);
  for (int i = ibegin; i < iend; ++i) drift(dt2, bs[i]);
  for (int i = ibegin; i < iend; ++i)
    for (int j = 0; j < i; ++j) kick(eta, dt, bs[i], bs[j]);
  if (nkicks) *nkicks += 1;
  for (int i = ibegin; i < iend; ++i) drift(dt2, bs[i]);
  advance_subsystem(dt2, eta, bs, ibegin, nkicks);
  sort_sub(bs, iend);
}

std::vector<LBody> load(int n, const int *id, const double *t, const double *dt_max,
                        const double *m, const double *q, const double *v) {
  std::vector<LBody> bs(n);
  for (int i = 0; i < n; ++i) {
    bs[i].id = id[i];
    bs[i].t = t[i];
    bs[i].dt_max = dt_max[i];
    bs[i].m = m[i];
    for (int k = 0; k < 3; ++k) {
      bs[i].q[k] = q[3 * i + k];
      bs[i].v[k] = v[3 * i + k];
    }
  }
  return bs;
}

void store(const std::vector<LBody> &bs, int *id, double *t, double *dt_max, double *m,
           double *q, double *v) {
  for (size_t i = 0; i < bs.size(); ++i) {
    id[i] = bs[i].id;
    t[i] = bs[i].t;
    dt_max[i] = bs[i].dt_max;
    m[i] = bs[i].m;
    for (int k = 0; k < 3; ++k) {
      q[3 * i + k] = bs[i].q[k];
      v[3 * i + k] = bs[i].v[k];
    }
  }
}

}  // namespace

extern "C" {

// Leapfrog.advance bs dt eta (leapfrog.ml:144-158).  Arrays in/out; output order is the
// reference's (sorted by dt_max), ids tell which body is which.
void oracle_leapfrog_advance(int n, int *id, double *t, double *dt_max, double *m, double *q,
                             double *v, double dt, double eta, long *nkicks) {
  auto bs = load(n, id, t, dt_max, m, q, v);
  for (auto &b : bs) b.dt_max = INFINITY;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) kick(eta, 0.0, bs[i], bs[j]);
  std::stable_sort(bs.begin(), bs.end(), [](const LBody &a, const LBody &b) {
    return ocaml_compare(a.dt_max, b.dt_max) < 0;
  });
  advance_subsystem(dt, eta, bs, n, nkicks);
  store(bs, id, t, dt_max, m, q, v);
}

// Leapfrog.advance_const_dt bs dt nsteps (leapfrog.ml:160-165).
void oracle_leapfrog_advance_const_dt(int n, int *id, double *t, double *dt_max, double *m,
                                      double *q, double *v, double dt, int nsteps) {
  auto bs = load(n, id, t, dt_max, m, q, v);
  for (int s = 0; s < nsteps; ++s) advance_subsystem(dt, INFINITY, bs, n, nullptr);
  store(bs, id, t, dt_max, m, q, v);
}

// Omelyan_advancer.OA.advance h bs (omelyan_advancer.ml:96-104) on m[n], q[3n], p[3n];
// *t += h.  q_star is internal scratch (overwritten from q at :90-92 before use).
void oracle_omelyan_advance(int n, const double *m, double *q, double *p, double *t, double h,
                            double eps2) {
  std::vector<double> qs(3 * (size_t)n);
  double gv[3];
  auto b_operate = [&]() {  // :44-56
    double f = h / 6.0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) {
        ob_grad_v(eps2, m[i], q + 3 * i, m[j], q + 3 * j, gv);
        for (int k = 0; k < 3; ++k) {
          p[3 * i + k] = p[3 * i + k] - f * gv[k];
          p[3 * j + k] = p[3 * j + k] + f * gv[k];
        }
      }
  };
  auto a_operate = [&]() {  // :58-65
    for (int i = 0; i < n; ++i) {
      double f = h / (2.0 * m[i]);
      for (int k = 0; k < 3; ++k) q[3 * i + k] = q[3 * i + k] + p[3 * i + k] * f;
    }
  };
  auto c_operate = [&]() {  // :67-94
    double pf = 2.0 * h / 3.0;
    for (size_t a = 0; a < qs.size(); ++a) qs[a] = q[a];
    for (int i = 0; i < n; ++i) {
      double f = (h * h) / (24.0 * m[i]);
      for (int j = i + 1; j < n; ++j) {
        double f2 = (h * h) / (24.0 * m[j]);
        ob_grad_v(eps2, m[i], q + 3 * i, m[j], q + 3 * j, gv);
        for (int k = 0; k < 3; ++k) {
          qs[3 * i + k] = qs[3 * i + k] - f * gv[k];
          qs[3 * j + k] = qs[3 * j + k] + f2 * gv[k];
        }
      }
    }
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) {
        ob_grad_v(eps2, m[i], qs.data() + 3 * i, m[j], qs.data() + 3 * j, gv);
        for (int k = 0; k < 3; ++k) {
          p[3 * i + k] = p[3 * i + k] - pf * gv[k];
          p[3 * j + k] = p[3 * j + k] + pf * gv[k];
        }
      }
  };
  b_operate();
  a_operate();
  c_operate();
  a_operate();
  b_operate();
  *t = *t + h;
}

}  // extern "C"
