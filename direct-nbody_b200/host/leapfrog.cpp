// leapfrog.cpp — host mirror of Leapfrog (leapfrog.ml:64-68, 106-165): drift, sort_sub,
// find_can_advance_index, advance_subsystem, advance, advance_const_dt.  The block recursion,
// drifts and sorts are host logic as in the reference; the two kick loops
// (leapfrog.ml:132-136, 151-155) are nbk_kick_block_sum calls (two rectangular kick sweeps).
//
// The kernel evaluates every pair of a sweep from the velocities at sweep start (SURVEY.md
// §3.3), so velocity kicks agree with the reference's sequential loop to rounding and dt_max
// to O(dt); with dt = 0 (the initial sweep) or eta = infinity (const-dt) they agree exactly.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.hpp"

namespace {

using namespace nbh;

// 0: host-buffer path (one nbk_kick_block_sum per kick block, host loop for const-dt steps);
// 1: bodies resident on the device, block recursion on the host (one launch group + one dt_max
//    read-back per kick block), const-dt steady state on the device;
// 2: as 1, and every subtree of advance_subsystem whose prefix fits runs on the device in one
//    launch (nbk_lf_advance_subsystem; agrees with 1 to rounding)
int g_leapfrog_resident = 2;
int g_leapfrog_tree_max = 0;  // 0: automatic
int g_leapfrog_device_sort_min = 512;  // resident sort_sub of at least this many rows: on the device

int lf_tree_max() {
  if (g_leapfrog_resident < 2) return 0;
  const int cap = nbk_lf_subsystem_max();
  if (g_leapfrog_tree_max > 0) return std::min(g_leapfrog_tree_max, cap);
  return cap;
}

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

// NBH_LF_PROFILE=1: wall-clock per phase of the host recursion, printed at the end of an advance
struct Prof {
  bool on = std::getenv("NBH_LF_PROFILE") != nullptr;
  double t[4] = {0, 0, 0, 0};  // kick blocks from the host, sorts, device subtrees, drifts
  long n[4] = {0, 0, 0, 0};
  std::chrono::steady_clock::time_point t0;
  void start() {
    if (on) t0 = std::chrono::steady_clock::now();
  }
  void stop(int k) {
    if (!on) return;
    t[k] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    n[k]++;
  }
  void report() const {
    if (!on) return;
    const char *name[4] = {"kick blocks (host recursion)", "sort_sub", "device subtrees", "drifts"};
    for (int k = 0; k < 4; ++k)
      std::fprintf(stderr, "  leapfrog %-30s %8ld calls %10.3f ms  %8.1f us/call\n", name[k], n[k],
                   1e3 * t[k], n[k] ? 1e6 * t[k] / n[k] : 0.0);
  }
};

struct Leap {
  nbk_ctx *ctx;
  Prof prof;
  std::vector<LBody> bs;
  std::vector<double> m, q, v, dv, tm;
  std::vector<int> perm;
  std::vector<char> small;
  long nkicks = 0;
  // resident: positions and velocities live on the device in bs order (nbk_lf_*); the host keeps
  // what the recursion needs (id, t, dt_max, m) and bs[i].q / .v are stale until fetch().
  bool resident = false;
  int tree_max = 0;  // resident only: prefixes up to this many bodies recurse on the device

  int to_device() {
    const int n = (int)bs.size();
    pack(n);
    tm.resize(n);
    for (int i = 0; i < n; ++i) tm[i] = bs[i].dt_max;
    NBH_TRY(ck(ctx, nbk_lf_upload(ctx, n, tm.data(), m.data(), q.data(), v.data())));
    resident = true;
    return 0;
  }
  int fetch() {
    const int n = (int)bs.size();
    q.resize(3 * (size_t)n);
    v.resize(3 * (size_t)n);
    NBH_TRY(ck(ctx, nbk_lf_download(ctx, nullptr, q.data(), v.data())));
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        bs[i].q[k] = q[3 * i + k];
        bs[i].v[k] = v[3 * i + k];
      }
    resident = false;
    return 0;
  }
  int begin_block(int ibegin, int iend) {  // leapfrog.ml:125-127
    for (int i = ibegin; i < iend; ++i) bs[i].dt_max = INFINITY;
    return 0;  // resident: nbk_lf_kick_block resets its active block itself
  }
  int drift_block(double dt, int ibegin, int iend) {  // leapfrog.ml:129-131, 137
    if (resident) {
      for (int i = ibegin; i < iend; ++i) bs[i].t = bs[i].t + dt;
      return ck(ctx, nbk_lf_drift(ctx, ibegin, iend, dt));
    }
    for (int i = ibegin; i < iend; ++i) drift(dt, bs[i]);
    return 0;
  }

  void drift(double dt, LBody &b) {  // leapfrog.ml:64-68
    b.t = b.t + dt;
    b.q[0] = b.q[0] + dt * b.v[0];
    b.q[1] = b.q[1] + dt * b.v[1];
    b.q[2] = b.q[2] + dt * b.v[2];
  }

  int sort_sub(int iend) {  // leapfrog.ml:106-111 (Array.fast_sort is stable)
    if (!resident) {
      std::stable_sort(bs.begin(), bs.begin() + iend, [](const LBody &a, const LBody &b) {
        return ocaml_compare(a.dt_max, b.dt_max) < 0;
      });
      return 0;
    }
    perm.resize(iend);
    if (iend >= g_leapfrog_device_sort_min) {
      // wide prefixes: the device ranks its own dt_max (the host's copy is the same values) and
      // moves the rows; only the permutation comes back
      NBH_TRY(ck(ctx, nbk_lf_sort_sub(ctx, iend, perm.data())));
      bool identity = true;
      for (int i = 0; i < iend && identity; ++i) identity = perm[i] == i;
      if (identity) return 0;
      // resident: q and v of the host records are stale until fetch(); move what the recursion
      // keeps per row (id, t, dt_max, m) only
      struct Small {
        int id;
        double t, dt_max, m;
      };
      small.resize(iend * sizeof(Small));
      Small *sm = reinterpret_cast<Small *>(small.data());
      for (int i = 0; i < iend; ++i) sm[i] = {bs[i].id, bs[i].t, bs[i].dt_max, bs[i].m};
      for (int i = 0; i < iend; ++i) {
        const Small &o = sm[perm[i]];
        bs[i].id = o.id;
        bs[i].t = o.t;
        bs[i].dt_max = o.dt_max;
        bs[i].m = o.m;
      }
      return 0;
    }
    for (int i = 0; i < iend; ++i) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) {
      return ocaml_compare(bs[a].dt_max, bs[b].dt_max) < 0;
    });
    bool identity = true;
    for (int i = 0; i < iend && identity; ++i) identity = perm[i] == i;
    if (identity) return 0;
    std::vector<LBody> tmp(bs.begin(), bs.begin() + iend);
    for (int i = 0; i < iend; ++i) bs[i] = tmp[perm[i]];
    return ck(ctx, nbk_lf_permute(ctx, iend, perm.data()));
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
  // drift_after: the half drift that follows the kick (leapfrog.ml:137), fused into the
  // resident call; the host path drifts in advance_subsystem.
  int kick_block(double eta, double dt, int ibegin, int iend, bool with_drift = false,
                 double drift_after = 0.0) {
    if (iend - ibegin <= 0) return 0;
    // eta = infinity makes every pair time-scale inf/NaN, which never lowers dt_max
    // (leapfrog.ml:94-104): skip that half of the kernel.
    const bool ts = !(std::isinf(eta) && eta > 0);
    if (resident) {
      tm.resize(iend);
      NBH_TRY(ck(ctx, nbk_lf_kick_block(ctx, iend, ibegin, eta, dt, ts ? 1 : 0, with_drift ? 1 : 0,
                                        drift_after,
                                        ts ? tm.data() : nullptr)));
      if (ts)
        for (int i = 0; i < iend; ++i) bs[i].dt_max = tm[i];
      if (with_drift)
        for (int i = ibegin; i < iend; ++i) bs[i].t = bs[i].t + drift_after;
      nkicks += 1;
      return 0;
    }
    pack(iend);
    dv.resize(3 * (size_t)iend);
    tm.resize(iend);
    NBH_TRY(ck(ctx, nbk_kick_block_sum(ctx, iend, ibegin, m.data(), q.data(), v.data(), eta, dt,
                                       dv.data(), ts ? tm.data() : nullptr)));
    apply(0, iend, ts);
    nkicks += 1;
    return 0;
  }

  // advance_subsystem for a prefix that fits the device-side recursion: one launch; the host's
  // records (id, m, t, dt_max) follow the returned permutation.
  int advance_subsystem_on_device(double dt, double eta, int iend) {
    tm.resize(iend);
    perm.resize(iend);
    std::vector<double> tt(iend);
    for (int i = 0; i < iend; ++i) tt[i] = bs[i].t;
    long long nk = 0;
    NBH_TRY(ck(ctx, nbk_lf_advance_subsystem(ctx, iend, dt, eta, tt.data(), tm.data(), perm.data(), &nk)));
    std::vector<LBody> tmp(bs.begin(), bs.begin() + iend);
    for (int i = 0; i < iend; ++i) {
      bs[i] = tmp[perm[i]];
      bs[i].t = tt[i];
      bs[i].dt_max = tm[i];
    }
    nkicks += (long)nk;
    return 0;
  }

  int advance_subsystem(double dt, double eta, int iend) {  // leapfrog.ml:119-142
    if (iend <= 0) return 0;
    if (resident && tree_max > 0 && iend <= tree_max) {
      prof.start();
      int rc = advance_subsystem_on_device(dt, eta, iend);
      prof.stop(2);
      return rc;
    }
    const int ibegin = find_can_advance_index(iend, dt);
    const double dt2 = dt / 2.0;
    NBH_TRY(begin_block(ibegin, iend));
    NBH_TRY(advance_subsystem(dt2, eta, ibegin));
    prof.start();
    NBH_TRY(drift_block(dt2, ibegin, iend));
    prof.stop(3);
    prof.start();
    if (resident) {
      NBH_TRY(kick_block(eta, dt, ibegin, iend, true, dt2));  // kick + second half drift, one call
    } else {
      NBH_TRY(kick_block(eta, dt, ibegin, iend));
      NBH_TRY(drift_block(dt2, ibegin, iend));
    }
    prof.stop(0);
    NBH_TRY(advance_subsystem(dt2, eta, ibegin));
    prof.start();
    int rc = sort_sub(iend);
    prof.stop(1);
    return rc;
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

void nbh_set_leapfrog_resident(int mode) { g_leapfrog_resident = mode; }
void nbh_set_leapfrog_tree_max(int bodies) { g_leapfrog_tree_max = bodies; }
void nbh_set_leapfrog_device_sort_min(int rows) { g_leapfrog_device_sort_min = rows; }

int nbh_leapfrog_advance(nbk_ctx *ctx, int n, int *id, double *t, double *dt_max, double *m,
                         double *q, double *v, double dt, double eta, long *nkicks) {
  Leap L;
  L.ctx = ctx;
  load(L, n, id, t, dt_max, m, q, v);  // Array.map copy (leapfrog.ml:145)
  for (auto &b : L.bs) b.dt_max = INFINITY;
  // the bodies stay on the device for the whole advance (nbk_lf_*); a multi-device context
  // has no resident Leapfrog state and goes through the host-buffer calls
  if (g_leapfrog_resident && n > 0 && nbk_device_count(ctx) == 1) {
    NBH_TRY(L.to_device());
    L.tree_max = lf_tree_max();
  }
  // "Kick" the entire system with dt = 0.0 to set up the time-scales (leapfrog.ml:146-155):
  // all pairs i<j == every body against every other.
  NBH_TRY(L.kick_block(eta, 0.0, 0, n));
  NBH_TRY(L.sort_sub(n));  // leapfrog.ml:156
  NBH_TRY(L.advance_subsystem(dt, eta, n));
  if (L.resident) NBH_TRY(L.fetch());
  store(L, id, t, dt_max, m, q, v);
  if (nkicks) *nkicks = L.nkicks;
  L.prof.report();
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
