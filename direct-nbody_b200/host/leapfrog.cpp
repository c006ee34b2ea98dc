// leapfrog.cpp — host mirror of Leapfrog (leapfrog.ml:64-68, 106-165): drift, sort_sub,
// find_can_advance_index, advance_subsystem, advance, advance_const_dt.  The block recursion,
// drifts and sorts are host logic as in the reference; the two kick loops
// (leapfrog.ml:132-136, 151-155) are nbk_kick_block_sum calls (two rectangular kick sweeps).
//
// The kernel evaluates every pair of a sweep from the velocities at sweep start (SURVEY.md
// §3.3), so velocity kicks agree with the reference's sequential loop to rounding and dt_max
// to O(dt); with dt = 0 (the initial sweep) or eta = infinity (const-dt) they agree exactly.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.hpp"

namespace {

using namespace nbh;

int g_leapfrog_resident = 1;  // const-dt steady state on the device (0: host loop, one sweep per step)

struct LBody {
  int id;
  double t, dt_max, m, q[3], v[3];
};

// OCaml's polymorphic compare on floats (leapfrog.ml:108, :156): nan below everything.
int ocaml_compare(double x, double y) {
  if (x < y) return -1;
  if (x > y) return 1;
  if (x == y) return 0;
  bool xn = std::isnan(x), yn = std::isnan(y);
  if (xn && yn) return 0;
  return xn ? -1 : 1;
}

struct Leap {
  nbk_ctx *ctx;
  std::vector<LBody> bs;
  std::vector<double> m, q, v, dv, tm;
  long nkicks = 0;

  void drift(double dt, LBody &b) {  // leapfrog.ml:64-68
    b.t = b.t + dt;
    b.q[0] = b.q[0] + dt * b.v[0];
    b.q[1] = b.q[1] + dt * b.v[1];
    b.q[2] = b.q[2] + dt * b.v[2];
  }

  void sort_sub(int iend) {  // leapfrog.ml:106-111 (Array.fast_sort is stable)
    std::stable_sort(bs.begin(), bs.begin() + iend, [](const LBody &a, const LBody &b) {
      return ocaml_compare(a.dt_max, b.dt_max) < 0;
    });
  }

  int find_can_advance_index(int iend, double dt) const {  // leapfrog.ml:113-117
    while (!(iend == 0 || bs[iend - 1].dt_max < dt)) --iend;
    return iend;
  }

  void pack(int iend) {
    m.resize(iend);
    q.resize(3 * (size_t)iend);
    v.resize(3 * (size_t)iend);
    for (int i = 0; i < iend; ++i) {
      m[i] = bs[i].m;
      for (int k = 0; k < 3; ++k) {
        q[3 * i + k] = bs[i].q[k];
        v[3 * i + k] = bs[i].v[k];
      }
    }
  }

  void apply(int lo, int hi, bool with_tscale) {
    for (int i = lo; i < hi; ++i) {
      for (int k = 0; k < 3; ++k) bs[i].v[k] = bs[i].v[k] + dv[3 * (size_t)(i - lo) + k];
      if (with_tscale && bs[i].dt_max > tm[i - lo]) bs[i].dt_max = tm[i - lo];
    }
  }

  // for i in [ibegin,iend) for j in [0,i): kick eta dt bs.(i) bs.(j)   (leapfrog.ml:132-136)
  // = targets [ibegin,iend) x sources [0,iend), and targets [0,ibegin) x sources [ibegin,iend).
  int kick_block(double eta, double dt, int ibegin, int iend) {
    if (iend - ibegin <= 0) return 0;
    pack(iend);
    // eta = infinity makes every pair time-scale inf/NaN, which never lowers dt_max
    // (leapfrog.ml:94-104): skip that half of the kernel.
    const bool ts = !(std::isinf(eta) && eta > 0);
    dv.resize(3 * (size_t)iend);
    tm.resize(iend);
    NBH_TRY(ck(ctx, nbk_kick_block_sum(ctx, iend, ibegin, m.data(), q.data(), v.data(), eta, dt,
                                       dv.data(), ts ? tm.data() : nullptr)));
    apply(0, iend, ts);
    nkicks += 1;
    return 0;
  }

  int advance_subsystem(double dt, double eta, int iend) {  // leapfrog.ml:119-142
    if (iend <= 0) return 0;
    const int ibegin = find_can_advance_index(iend, dt);
    const double dt2 = dt / 2.0;
    for (int i = ibegin; i < iend; ++i) bs[i].dt_max = INFINITY;
    NBH_TRY(advance_subsystem(dt2, eta, ibegin));
    for (int i = ibegin; i < iend; ++i) drift(dt2, bs[i]);
    NBH_TRY(kick_block(eta, dt, ibegin, iend));
    for (int i = ibegin; i < iend; ++i) drift(dt2, bs[i]);
    NBH_TRY(advance_subsystem(dt2, eta, ibegin));
    sort_sub(iend);
    return 0;
  }
};

void load(Leap &L, int n, const int *id, const double *t, const double *dt_max, const double *m,
          const double *q, const double *v) {
  L.bs.resize(n);
  for (int i = 0; i < n; ++i) {
    L.bs[i].id = id[i];
    L.bs[i].t = t[i];
    L.bs[i].dt_max = dt_max[i];
    L.bs[i].m = m[i];
    for (int k = 0; k < 3; ++k) {
      L.bs[i].q[k] = q[3 * i + k];
      L.bs[i].v[k] = v[3 * i + k];
    }
  }
}

void store(const Leap &L, int *id, double *t, double *dt_max, double *m, double *q, double *v) {
  for (size_t i = 0; i < L.bs.size(); ++i) {
    id[i] = L.bs[i].id;
    t[i] = L.bs[i].t;
    dt_max[i] = L.bs[i].dt_max;
    m[i] = L.bs[i].m;
    for (int k = 0; k < 3; ++k) {
      q[3 * i + k] = L.bs[i].q[k];
      v[3 * i + k] = L.bs[i].v[k];
    }
  }
}

}  // namespace

extern "C" {

void nbh_set_leapfrog_resident(int on) { g_leapfrog_resident = on; }

int nbh_leapfrog_advance(nbk_ctx *ctx, int n, int *id, double *t, double *dt_max, double *m,
                         double *q, double *v, double dt, double eta, long *nkicks) {
  Leap L;
  L.ctx = ctx;
  load(L, n, id, t, dt_max, m, q, v);  // Array.map copy (leapfrog.ml:145)
  for (auto &b : L.bs) b.dt_max = INFINITY;
  // "Kick" the entire system with dt = 0.0 to set up the time-scales (leapfrog.ml:146-155):
  // all pairs i<j == every body against every other.
  NBH_TRY(L.kick_block(eta, 0.0, 0, n));
  std::stable_sort(L.bs.begin(), L.bs.end(), [](const LBody &a, const LBody &b) {
    return ocaml_compare(a.dt_max, b.dt_max) < 0;
  });
  NBH_TRY(L.advance_subsystem(dt, eta, n));
  store(L, id, t, dt_max, m, q, v);
  if (nkicks) *nkicks = L.nkicks;
  return 0;
}

int nbh_leapfrog_advance_const_dt(nbk_ctx *ctx, int n, int *id, double *t, double *dt_max,
                                  double *m, double *q, double *v, double dt, int nsteps) {
  Leap L;
  L.ctx = ctx;
  load(L, n, id, t, dt_max, m, q, v);
  for (int s = 0; s < nsteps; ++s) {
    // Steady state: every body has dt_max >= dt, so find_can_advance_index returns 0, the
    // recursion is empty and each remaining call is one plain DKD step over all bodies with
    // dt_max left at infinity and the order untouched (leapfrog.ml:119-142).  Run all of them
    // on the device without coming back to the host.
    bool steady = true;
    for (const LBody &b : L.bs) steady = steady && !(b.dt_max < dt);
    if (steady && g_leapfrog_resident) {
      const int rest = nsteps - s;
      L.pack(n);
      NBH_TRY(ck(ctx, nbk_bodies_upload(ctx, n, L.m.data(), L.q.data(), L.v.data(), 1)));
      NBH_TRY(ck(ctx, nbk_resident_leapfrog_steps(ctx, dt, rest, L.q.data(), L.v.data())));
      const double dt2 = dt / 2.0;
      for (int i = 0; i < n; ++i) {
        LBody &b = L.bs[i];
        for (int k = 0; k < 3; ++k) {
          b.q[k] = L.q[3 * i + k];
          b.v[k] = L.v[3 * i + k];
        }
        for (int r = 0; r < rest; ++r) b.t = (b.t + dt2) + dt2;  // two drifts per step
        b.dt_max = INFINITY;
      }
      break;
    }
    NBH_TRY(L.advance_subsystem(dt, INFINITY, n));
  }
  store(L, id, t, dt_max, m, q, v);
  return 0;
}

}  // extern "C"
