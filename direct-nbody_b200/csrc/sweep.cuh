// sweep.cuh — the one sweep engine every O(N^2) pair loop of the path runs on (sm_100a only).
//
// Shape: targets x sources is cut into units of TI targets x TJ sources.
// Units are numbered i-tile-major and dealt to a PERSISTENT grid of G = (#SMs x resident
// CTAs/SM) CTAs in equal contiguous runs ("stream-K" over the source axis), so the grid is
// always full whatever N is and a handful of targets still spreads its sources over the whole
// chip.  Inside a CTA each thread owns IPT targets in registers (i-parallel); source tiles
// (TJ packed bodies = TJ*64 B, contiguous) are staged into shared memory by the TMA engine
// (cp.async.bulk + mbarrier complete_tx; SASS: UBLKCP / SYNCS) through a STAGES-deep
// full/empty ring, and every lane reads the same source with broadcast LDS.128.
// Accumulators stay in registers for a whole run of source tiles.  A run that covers a
// complete source range writes final results; a partial run writes its accumulators to a
// per-CTA slot and a small fix-up kernel combines the slots of each i-tile in ascending
// source order - fixed order, so results are reproducible run to run (no atomics).
//
// No tensor cores: a pair interaction is not a contraction.  The bound is the FP64 pipe.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nbk {

constexpr int PK = 8;            // doubles per packed body {x,y,z,m,vx,vy,vz,0}
constexpr int TJ = 128;          // sources per tile (8 KB per stage)
constexpr int STAGES = 3;
constexpr int MAX_PEERS = 8;      // GPUs of one NVSwitch box

// ---- PTX helpers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NBK_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NBK_DONE_%=;\n"
      "bra NBK_WAIT_%=;\n"
      "NBK_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// System-scope acquire load of a flag another GPU writes over NVLink.
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Orders the generic-proxy acquire above before the async-proxy (TMA) reads that follow.
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async;" ::: "memory");
}

// 1/sqrt(x): MUFU.RSQ64H seed (~2^-22 relative) + one third-order step
// y = y0 (1 + e/2 + 3 e^2/8), e = 1 - x y0^2: error ~ (5/16) e^3 < 2^-64, so the result is
// the correctly rounded value to within the last two roundings.  2 DMUL + 3 DFMA.  No
// denormal/huge-exponent slow path: r^2 of a physical configuration is a normal number, and
// x = 0 (two distinct coincident bodies, eps2 = 0) yields NaN - the reference gives inf/NaN
// there too (base.ml:77-80).
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  double t = x * y0;
  double e = fma(-t, y0, 1.0);
  double p = fma(0.375, e, 0.5);
  double q = e * y0;
  return fma(p, q, y0);
}

// 1/x: MUFU.RCP64H seed + two Newton steps (4 DFMA), ~1 ulp.  For places that need speed and
// not the correctly rounded quotient (the device-resident Hermite predictor).
__device__ __forceinline__ double rcp_nr(double x) {
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  double e = fma(-x, y0, 1.0);
  y0 = fma(y0, e, y0);
  e = fma(-x, y0, 1.0);
  return fma(y0, e, y0);
}

// Grid-wide barrier of a cooperative launch: every CTA adds 1 to a counter in L2 and waits for
// it to reach `target` (G x the number of barriers so far).
__device__ __forceinline__ void adv_grid_barrier(unsigned long long *bar, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1ULL);
    unsigned long long v;
    const long long c0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (clock64() - c0 > (20LL << 30)) __trap();  // ~10 s: a CTA is missing, do not hang the GPU
    }
    __threadfence();
  }
  __syncthreads();
}

// ---- sweep parameters ----------------------------------------------------------------------
struct SweepParams {
  const double *packed;  // [n_total][PK]
  const double *aux;     // second packed array (tangent directions / unused)
  int aux_stride;        // doubles between successive directions in aux
  int K;                 // tangent directions
  int t0, t1, s0, s1;
  int n_it, n_jt;             // tiles along targets / sources
  long long units;            // n_it * n_jt
  long long units_per_cta;    // ceil(units / G)
  double eps2, dt, eta;
  double *partial;  // [2*G][NACC][TI]
  double *out[4];   // op-specific outputs (device)
  int accumulate;
  // Sharded sources (sweep_kernel<..., PEER = true> only): body j lives in row j % shard_rows of
  // shard[j / shard_rows], a peer-mapped array on the GPU that owns it (NVLink P2P).  `packed`
  // is then this GPU's own shard, biased so that packed + i*PK is target i.  ready != NULL:
  // rank r's rows may be read once ready[r] >= epoch (peers store their flag with release.sys
  // after packing, nbk::publish_kernel).
  const double *shard[MAX_PEERS];
  const double *aux_shard[MAX_PEERS];
  int shard_rows;
  int self_rank;
  const unsigned long long *ready;
  unsigned long long epoch;
  unsigned long long wait_ns;  // how long a CTA waits for a peer's flag before trapping; 0 = forever
  // OpKick<true>: per target (index i - t0), the smallest exact pair time-scale any CTA has found
  // so far, as the bits of a non-negative double (integer order = numeric order); +inf at launch.
  // NULL = not shared (each run of sources warms its filter up on its own).
  unsigned long long *tbest;
};

// ---- pair operations -----------------------------------------------------------------------
// Each op: NACC accumulators per target, Tgt = the target's registers, pair() = one
// interaction with the source held in (a,b,c,d) = packed body as 4 double2, combine() for
// the fix-up, store() = final write with the reference's sign/scale convention.

// Base.grad_v summed (base.ml:73-80): acc = sum m_j (q_j - q_i) / s^3; gsum_i = -m_i acc.
struct OpGradV {
  static constexpr int NACC = 3;
  static constexpr int ND2 = 2;  // double2 loads per source: (x,y) (z,m)
  struct Tgt {
    double x, y, z;
  };
  __device__ __forceinline__ static void load(Tgt &t, const double *pk, const SweepParams &, int) {
    t.x = pk[0];
    t.y = pk[1];
    t.z = pk[2];
  }
  __device__ __forceinline__ static void zero(double *acc) { acc[0] = acc[1] = acc[2] = 0.0; }
  template <bool CHECK>
  __device__ __forceinline__ static void pair(const Tgt &t, double *acc, const double2 *sj,
                                              const double2 *, bool self, const SweepParams &P) {
    double2 a = sj[0], b = sj[1];
    double dx = a.x - t.x, dy = a.y - t.y, dz = b.x - t.z;
    double r2 = fma(dx, dx, P.eps2);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    double mj = b.y;
    if (CHECK) {
      r2 = self ? 1.0 : r2;
      mj = self ? 0.0 : mj;
    }
    double y = rsqrt_nr(r2);
    double w = (mj * y) * (y * y);
    acc[0] = fma(w, dx, acc[0]);
    acc[1] = fma(w, dy, acc[1]);
    acc[2] = fma(w, dz, acc[2]);
  }
  __device__ __forceinline__ static void combine(double *a, const double *b) {
    a[0] += b[0];
    a[1] += b[1];
    a[2] += b[2];
  }
  __device__ __forceinline__ static void store(const SweepParams &P, int ig, const double *acc) {
    double s = -P.packed[(size_t)ig * PK + 3];
    double *o = P.out[0] + 3 * (size_t)(ig - P.t0);
    for (int k = 0; k < 3; ++k) o[k] = P.accumulate ? o[k] + s * acc[k] : s * acc[k];
  }
};

// Base.grad_v_dgrad_v summed (base.ml:91-110).  With d = q_j - q_i, u = v_j - v_i, v = p/m:
//   acc  = sum m_j d / s^3                         gsum_i  = -m_i acc
//   jerk = sum m_j (u - 3 (d.u) d / s^2) / s^3     dgsum_i = -m_i jerk
// 6 DADD + 8 DMUL + 17 DFMA + 1 MUFU per interaction.
struct OpGradVDGradV {
  static constexpr int NACC = 6;
  static constexpr int ND2 = 4;
  struct Tgt {
    double x, y, z, vx, vy, vz;
  };
  __device__ __forceinline__ static void load(Tgt &t, const double *pk, const SweepParams &, int) {
    t.x = pk[0];
    t.y = pk[1];
    t.z = pk[2];
    t.vx = pk[4];
    t.vy = pk[5];
    t.vz = pk[6];
  }
  __device__ __forceinline__ static void zero(double *acc) {
    for (int k = 0; k < 6; ++k) acc[k] = 0.0;
  }
  template <bool CHECK>
  __device__ __forceinline__ static void pair(const Tgt &t, double *acc, const double2 *sj,
                                              const double2 *, bool self, const SweepParams &P) {
    double2 a = sj[0], b = sj[1], c = sj[2], d = sj[3];
    double dx = a.x - t.x, dy = a.y - t.y, dz = b.x - t.z;
    double ux = c.x - t.vx, uy = c.y - t.vy, uz = d.x - t.vz;
    double r2 = fma(dx, dx, P.eps2);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    double rv = dx * ux;
    rv = fma(dy, uy, rv);
    rv = fma(dz, uz, rv);
    double mj = b.y;
    if (CHECK) {
      r2 = self ? 1.0 : r2;
      mj = self ? 0.0 : mj;
    }
    double y = rsqrt_nr(r2);
    double y2 = y * y;
    double w = (mj * y) * y2;
    double a3 = (rv * y2) * -3.0;
    acc[0] = fma(w, dx, acc[0]);
    acc[1] = fma(w, dy, acc[1]);
    acc[2] = fma(w, dz, acc[2]);
    double tx = fma(a3, dx, ux), ty = fma(a3, dy, uy), tz = fma(a3, dz, uz);
    acc[3] = fma(w, tx, acc[3]);
    acc[4] = fma(w, ty, acc[4]);
    acc[5] = fma(w, tz, acc[5]);
  }
  __device__ __forceinline__ static void combine(double *a, const double *b) {
    for (int k = 0; k < 6; ++k) a[k] += b[k];
  }
  __device__ __forceinline__ static void store(const SweepParams &P, int ig, const double *acc) {
    double s = -P.packed[(size_t)ig * PK + 3];
    double *g = P.out[0] + 3 * (size_t)(ig - P.t0);
    double *dg = P.out[1] + 3 * (size_t)(ig - P.t0);
    for (int k = 0; k < 3; ++k) {
      g[k] = P.accumulate ? g[k] + s * acc[k] : s * acc[k];
      dg[k] = P.accumulate ? dg[k] + s * acc[3 + k] : s * acc[3 + k];
    }
  }
};

// Base.v summed (base.ml:60-63): acc = sum m_j / s; phi_i = -m_i acc.
struct OpV {
  static constexpr int NACC = 1;
  static constexpr int ND2 = 2;
  struct Tgt {
    double x, y, z;
  };
  __device__ __forceinline__ static void load(Tgt &t, const double *pk, const SweepParams &, int) {
    t.x = pk[0];
    t.y = pk[1];
    t.z = pk[2];
  }
  __device__ __forceinline__ static void zero(double *acc) { acc[0] = 0.0; }
  template <bool CHECK>
  __device__ __forceinline__ static void pair(const Tgt &t, double *acc, const double2 *sj,
                                              const double2 *, bool self, const SweepParams &P) {
    double2 a = sj[0], b = sj[1];
    double dx = a.x - t.x, dy = a.y - t.y, dz = b.x - t.z;
    double r2 = fma(dx, dx, P.eps2);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    double mj = b.y;
    if (CHECK) {
      r2 = self ? 1.0 : r2;
      mj = self ? 0.0 : mj;
    }
    acc[0] = fma(mj, rsqrt_nr(r2), acc[0]);
  }
  __device__ __forceinline__ static void combine(double *a, const double *b) { a[0] += b[0]; }
  __device__ __forceinline__ static void store(const SweepParams &P, int ig, const double *acc) {
    double s = -P.packed[(size_t)ig * PK + 3];
    double *o = P.out[0] + (size_t)(ig - P.t0);
    o[0] = P.accumulate ? o[0] + s * acc[0] : s * acc[0];
  }
};

// Leapfrog.kick summed (leapfrog.ml:70-104), pre-kick velocities.  The velocity half uses
// the fast rsqrt (dv_i = dt * sum m_j d / r^3, unsoftened).  The time-scale half must return,
// bit for bit, the minimum over sources of the reference's pair time-scale, which costs 2 sqrt
// + 5 div in IEEE arithmetic - far too much to run on every pair.  So every pair goes through a
// FILTER: a sufficient condition, free of divisions, square roots and d.u, for "this pair's
// time-scale exceeds the running minimum by more than 0.2 %".  Only pairs that fail it - about
// ln N per target, plus anything NaN - run the reference's exact operation sequence (correctly
// rounded sqrt/div, no contraction), so the minimum is bit-identical to the reference formula.
//
// Filter (FP64, on values the velocity half has already: r2, y = 1/r; plus |u|^2).  With
// M = m_i + m_j, s = d.u / r^2 (leapfrog.ml:94-101):
//   tff0 = eta sqrt(r^3/M),  tff = tff0 / |1 - a|,  a = 0.75 s tff0
//   tc0  = eta r/|u|,        tc  = tc0  / |1 - b|,  b = 0.5 s tc0 (1 + z),  z = M/(|u|^2 r)
// Cauchy-Schwarz gives |s| <= |u|/r = eta/tc0, hence |b| <= 0.5 eta (1 + z) and
// |a| <= 0.75 eta tff0/tc0, and tff0/tc0 = 1/sqrt(z).  If z <= 1 (the pair moves faster than its
// mutual orbital speed - all but bound pairs) then tff0 >= tc0, tc >= tc0/(1 + eta) and
// tff >= tff0/(1 + 0.75 eta tff0/tc0) >= tc0/(1 + 0.75 eta): the pair's time-scale is at least
// tc0/(1 + eta).  So the pair is "far" when
//   |M| y <= |u|^2            (z <= 1)                                 and
//   c r^2  >  |u|^2           (tc0 > T (1 + eta)),  c = (eta / (T (1 + eta)))^2, T = 1.002 min-so-far
// 1 DADD + 2 DMUL + 2 DSETP on top of |u|^2, no conversions, no FP32 work (the earlier FP32-pipe
// filter cost 4 F2F.F32.F64 on the XU pipe and ~30 FP32 issue slots per pair: 50 ms at
// N = 131072; an FP64 instruction holds the dispatch port for >= 2 cycles, so every non-FP64
// instruction adds to the FP64 time instead of hiding under it).  NaN anywhere fails a
// comparison and takes the exact path; a zero or negative total mass leaves tff NaN or
// infinite in the reference too, where the time-scale is tc, and the bound on tc holds for
// |z| <= 1.  State per target (acc[4]) = c; 0 (never "far") until a first exact value exists,
// for eta <= 0 or NaN, or whenever T leaves [1e-100, 1e100].
template <bool WITH_TSCALE> struct OpKick {
  // acc: [0..2] velocity sums; [3] the minimum exact time-scale this run has evaluated; [4] the
  // filter constant c and [5] the bound T it was derived from: min([3], what other runs of the
  // same target have published, SweepParams::tbest)
  static constexpr int NACC = WITH_TSCALE ? 6 : 3;
  static constexpr int ND2 = WITH_TSCALE ? 4 : 2;
  struct Tgt {
    double x, y, z, vx, vy, vz, m;
  };
  __device__ __forceinline__ static void load(Tgt &t, const double *pk, const SweepParams &, int) {
    t.x = pk[0];
    t.y = pk[1];
    t.z = pk[2];
    t.m = pk[3];
    t.vx = pk[4];
    t.vy = pk[5];
    t.vz = pk[6];
  }
  __device__ __forceinline__ static double filter_state(double tmin, double eta) {
    const double T = tmin * 1.002;
    double c = 0.0;
    if (T > 1e-100 && T < 1e100 && eta > 0.0 && eta < 1e100) {
      const double k = eta / (T * (1.0 + eta));
      c = (k * k) * 0.999999;
    }
    return c;
  }
  __device__ __forceinline__ static void zero(double *acc) {
    acc[0] = acc[1] = acc[2] = 0.0;
    if (WITH_TSCALE) {
      acc[3] = __longlong_as_double(0x7ff0000000000000LL);
      acc[4] = 0.0;
      acc[5] = __longlong_as_double(0x7ff0000000000000LL);
    }
  }
  // A new exact value for this target: the minimum, and the filter if it is the tightest bound.
  __device__ __forceinline__ static void record(double *acc, double ts, double eta) {
    if (acc[3] > ts) {  // NaN ts leaves acc[3] alone, leapfrog.ml:103
      acc[3] = ts;
      if (ts < acc[5]) {
        acc[5] = ts;
        acc[4] = filter_state(ts, eta);
      }
    }
  }
  // Between source tiles: the sources of a target are split into RUNS over several CTAs (stream-K,
  // and one process per GPU shards them further), and every run would warm its filter up from
  // +infinity on its own - with 128 (lane, target) pairs per warp the first ~128 (1 + ln(K / 128))
  // of a run's K iterations take the exact branch: 0.4 ms per run, 25 % of an 8-GPU kick sweep.
  // So the runs of a target SHARE their bound: each publishes its minimum (atomicMin on the bits)
  // and adopts a smaller one.  Any published value is the exact time-scale of a real pair of this
  // target, hence an upper bound of its final minimum: the filter stays a sufficient condition,
  // and the run that holds the true minimum pair still evaluates it - the combined result is
  // bit-identical, whatever the timing.
  static constexpr int TILE_HOOK = 1;
  template <int IPT>
  __device__ __forceinline__ static void tile_end(double (&acc)[IPT][NACC], const int (&ig)[IPT],
                                                  const SweepParams &P) {
    if constexpr (WITH_TSCALE) {
      if (P.tbest == nullptr) return;
#pragma unroll
      for (int r = 0; r < IPT; ++r) {
        if (ig[r] >= P.t1) continue;
        unsigned long long *slot = P.tbest + (ig[r] - P.t0);
        const double g = __longlong_as_double((long long)__ldcg(slot));
        if (acc[r][3] < g) {
          atomicMin(slot, (unsigned long long)__double_as_longlong(acc[r][3]));
        } else if (g < acc[r][5]) {
          acc[r][5] = g;
          acc[r][4] = filter_state(g, P.eta);
        }
      }
    }
  }
  // The reference's exact pair time-scale (leapfrog.ml:80-104, unfused, correctly rounded).
  __device__ __noinline__ static double exact_tscale(double dx, double dy, double dz, double ux,
                                                     double uy, double uz, double mi, double mj,
                                                     double eta) {
    const double r2e = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double vsqe = __dadd_rn(__dadd_rn(__dmul_rn(ux, ux), __dmul_rn(uy, uy)), __dmul_rn(uz, uz));
    const double vdre = __dadd_rn(__dadd_rn(__dmul_rn(dx, ux), __dmul_rn(dy, uy)), __dmul_rn(dz, uz));
    const double mtot = __dadd_rn(mi, mj);
    double r = __dsqrt_rn(r2e);
    double r3 = __dmul_rn(r2e, r);
    double tff = __dmul_rn(eta, __dsqrt_rn(__ddiv_rn(r3, mtot)));
    double tc = __dmul_rn(eta, __dsqrt_rn(__ddiv_rn(r2e, vsqe)));
    double svdr = __ddiv_rn(vdre, r2e);
    double dtff = __dmul_rn(__dmul_rn(1.5, svdr), tff);
    double dtc = __dmul_rn(__dmul_rn(svdr, tc), __dadd_rn(1.0, __ddiv_rn(mtot, __dmul_rn(vsqe, r))));
    tff = fabs(__ddiv_rn(tff, __dsub_rn(1.0, __dmul_rn(0.5, dtff))));
    tc = fabs(__ddiv_rn(tc, __dsub_rn(1.0, __dmul_rn(0.5, dtc))));
    return (tff < tc) ? tff : tc;
  }
  // true = this pair's time-scale certainly exceeds 1.002 x the running minimum.
  __device__ __forceinline__ static bool far_pair(double r2, double y, double vsq, double M,
                                                  double c) {
    return (fabs(M) * y <= vsq) & (c * r2 > vsq);
  }
  template <bool CHECK>
  __device__ __forceinline__ static void pair(const Tgt &t, double *acc, const double2 *sj,
                                              const double2 *, bool self, const SweepParams &P) {
    double2 a = sj[0], b = sj[1];
    double dx = a.x - t.x, dy = a.y - t.y, dz = b.x - t.z;
    double r2 = dx * dx;
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    double mj = b.y;
    if (CHECK) {
      r2 = self ? 1.0 : r2;
      mj = self ? 0.0 : mj;
    }
    const double y = rsqrt_nr(r2);
    const double w = (mj * y) * (y * y);
    acc[0] = fma(w, dx, acc[0]);
    acc[1] = fma(w, dy, acc[1]);
    acc[2] = fma(w, dz, acc[2]);
    if (WITH_TSCALE) {
      double2 c = sj[2], d = sj[3];
      const double ux = c.x - t.vx, uy = c.y - t.vy, uz = d.x - t.vz;
      double vsq = ux * ux;
      vsq = fma(uy, uy, vsq);
      vsq = fma(uz, uz, vsq);
      const bool far = far_pair(r2, y, vsq, t.m + b.y, acc[4]);
      const bool skip = CHECK ? self : false;
      if (!far && !skip) {
        record(acc, exact_tscale(dx, dy, dz, ux, uy, uz, t.m, b.y, P.eta), P.eta);
      }
    }
  }
  // One source against the IPT targets of a thread, phase by phase: the IPT filters run as
  // independent instruction streams and ONE branch guards the (rare) exact evaluations.
  static constexpr int BATCHED = 1;
  template <bool CHECK, int IPT>
  __device__ __forceinline__ static void pairs(const Tgt (&t)[IPT], double (&acc)[IPT][NACC],
                                               const double2 *sj, int jg, const int (&ig)[IPT],
                                               const SweepParams &P) {
    if constexpr (!WITH_TSCALE) {
#pragma unroll
      for (int r = 0; r < IPT; ++r) pair<CHECK>(t[r], acc[r], sj, nullptr, jg == ig[r], P);
    } else {
      const double2 a = sj[0], b = sj[1], c = sj[2], d = sj[3];
      double dx[IPT], dy[IPT], dz[IPT], ux[IPT], uy[IPT], uz[IPT], r2[IPT], mj[IPT], y0[IPT];
      double y[IPT], vsq[IPT];
#pragma unroll
      for (int r = 0; r < IPT; ++r) {
        dx[r] = a.x - t[r].x;
        dy[r] = a.y - t[r].y;
        dz[r] = b.x - t[r].z;
      }
#pragma unroll
      for (int r = 0; r < IPT; ++r) {
        r2[r] = dx[r] * dx[r];
        r2[r] = fma(dy[r], dy[r], r2[r]);
        r2[r] = fma(dz[r], dz[r], r2[r]);
        mj[r] = b.y;
        if (CHECK) {
          const bool self = jg == ig[r];
          r2[r] = self ? 1.0 : r2[r];
          mj[r] = self ? 0.0 : mj[r];
        }
      }
#pragma unroll
      for (int r = 0; r < IPT; ++r) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[r]) : "d"(r2[r]));
#pragma unroll
      for (int r = 0; r < IPT; ++r) {
        ux[r] = c.x - t[r].vx;
        uy[r] = c.y - t[r].vy;
        uz[r] = d.x - t[r].vz;
      }
#pragma unroll
      for (int r = 0; r < IPT; ++r) {
        vsq[r] = ux[r] * ux[r];
        vsq[r] = fma(uy[r], uy[r], vsq[r]);
        vsq[r] = fma(uz[r], uz[r], vsq[r]);
      }
#pragma unroll
      for (int r = 0; r < IPT; ++r) {
        double tt = r2[r] * y0[r];
        double e = fma(-tt, y0[r], 1.0);
        double pp = fma(0.375, e, 0.5);
        double q = e * pp;
        y[r] = fma(y0[r], q, y0[r]);
      }
      // hot path: ONE predicate "every target's pair is far" (2 DSETP per target, chained); the
      // cold path re-derives which targets need the exact sequence
      double My[IPT], cr2[IPT];
      bool allfar = true;
#pragma unroll
      for (int r = 0; r < IPT; ++r) {
        My[r] = fabs(t[r].m + b.y) * y[r];
        cr2[r] = acc[r][4] * r2[r];
        const bool self = CHECK ? (jg == ig[r]) : false;
        allfar = allfar && (((My[r] <= vsq[r]) && (cr2[r] > vsq[r])) || self);
      }
#pragma unroll
      for (int r = 0; r < IPT; ++r) {
        const double w = (mj[r] * y[r]) * (y[r] * y[r]);
        acc[r][0] = fma(w, dx[r], acc[r][0]);
        acc[r][1] = fma(w, dy[r], acc[r][1]);
        acc[r][2] = fma(w, dz[r], acc[r][2]);
      }
      if (!allfar) {
        // Which of this lane's targets need the exact sequence; then one exact evaluation per
        // trip of a lane-local loop, each lane working on ITS next target.  The warp makes as
        // many trips as its busiest lane has needs - usually one - where a loop over r in
        // lockstep made one trip for every r that ANY lane needed (four, while a run is young).
        unsigned need = 0;
#pragma unroll
        for (int r = 0; r < IPT; ++r) {
          const bool self = CHECK ? (jg == ig[r]) : false;
          const bool far = (My[r] <= vsq[r]) && (cr2[r] > vsq[r]);
          if (!far && !self) need |= 1u << r;
        }
        while (need) {
          const int r = __ffs(need) - 1;
          need &= need - 1;
          double ex = dx[0], ey = dy[0], ez = dz[0], fx = ux[0], fy = uy[0], fz = uz[0], mi = t[0].m;
#pragma unroll
          for (int rr = 1; rr < IPT; ++rr)
            if (rr == r) {
              ex = dx[rr], ey = dy[rr], ez = dz[rr];
              fx = ux[rr], fy = uy[rr], fz = uz[rr];
              mi = t[rr].m;
            }
          const double ts = exact_tscale(ex, ey, ez, fx, fy, fz, mi, b.y, P.eta);
#pragma unroll
          for (int rr = 0; rr < IPT; ++rr)
            if (rr == r) record(acc[rr], ts, P.eta);
        }
      }
    }
  }
  __device__ __forceinline__ static void combine(double *a, const double *b) {
    a[0] += b[0];
    a[1] += b[1];
    a[2] += b[2];
    if (WITH_TSCALE)
      if (a[3] > b[3]) a[3] = b[3];
  }
  __device__ __forceinline__ static void store(const SweepParams &P, int ig, const double *acc) {
    double *o = P.out[0] + 3 * (size_t)(ig - P.t0);
    for (int k = 0; k < 3; ++k) o[k] = P.dt * acc[k];
    if (WITH_TSCALE) P.out[1][ig - P.t0] = acc[3];
  }
};

// Analysis.binary_energy scanned over sources (analysis.ml:809-822, the pair function of
// Analysis.binaries / binaries_threshold / tightest_binary, analysis.ml:836-876):
//   e(i,j) = 0.5 mu |v_j - v_i|^2 + Base.v(i,j),  mu = m_i m_j / (m_i + m_j)
// evaluated with the reference's IEEE operation sequence (no contraction), so the decisions
// e < threshold and the minimum are bit-exact.  Per target: acc[0] = min_j e(i,j),
// acc[1] = the partner index attaining it (lowest index on ties), acc[2] = #{j : e < thr}.
// P.dt carries the threshold.
struct OpBinary {
  static constexpr int NACC = 3;
  static constexpr int ND2 = 4;
  struct Tgt {
    double x, y, z, vx, vy, vz, m;
  };
  __device__ __forceinline__ static void load(Tgt &t, const double *pk, const SweepParams &, int) {
    t.x = pk[0];
    t.y = pk[1];
    t.z = pk[2];
    t.m = pk[3];
    t.vx = pk[4];
    t.vy = pk[5];
    t.vz = pk[6];
  }
  __device__ __forceinline__ static void zero(double *acc) {
    acc[0] = __longlong_as_double(0x7ff0000000000000LL);
    acc[1] = -1.0;
    acc[2] = 0.0;
  }
  template <bool CHECK>
  __device__ __forceinline__ static void pair(const Tgt &t, double *acc, const double2 *sj,
                                              const double2 *, bool self, const SweepParams &P) {
    if (self) return;
    const double2 a = sj[0], b = sj[1], c = sj[2], d = sj[3];
    const double ux = __dsub_rn(c.x, t.vx), uy = __dsub_rn(c.y, t.vy), uz = __dsub_rn(d.x, t.vz);
    const double dx = __dsub_rn(t.x, a.x), dy = __dsub_rn(t.y, a.y), dz = __dsub_rn(t.z, b.x);
    const double mj = b.y;
    const double mu = __ddiv_rn(__dmul_rn(t.m, mj), __dadd_rn(t.m, mj));
    const double v2 = __dadd_rn(__dadd_rn(__dmul_rn(ux, ux), __dmul_rn(uy, uy)), __dmul_rn(uz, uz));
    const double ke = __dmul_rn(__dmul_rn(0.5, mu), v2);
    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double r = __dsqrt_rn(__dadd_rn(r2, P.eps2));
    const double pe = __ddiv_rn(__dmul_rn(-t.m, mj), r);
    const double e = __dadd_rn(ke, pe);
    // the packed source carries its global index in slot 7 (see nbk_binary_scan)
    const double jidx = d.y;
    if (e < acc[0] || (e == acc[0] && jidx < acc[1])) {
      acc[0] = e;
      acc[1] = jidx;
    }
    if (e < P.dt) acc[2] += 1.0;
  }
  // One source against the IPT targets of a thread, with a FILTER in front of the exact sequence
  // (2 IEEE divisions + 1 square root per pair, which made the scan 4x slower than the acc+jerk
  // sweep).  A pair matters only if e < thr (it is counted) or e <= the running minimum (it may
  // become the partner); with T = max(thr, emin), M = m_i + m_j > 0, y = 1/sqrt(r^2 + eps2):
  //   e > T  <=>  m_i m_j (v^2/2 - M y) > T M,
  // evaluated by fast arithmetic (FMA, rsqrt with one third-order step) with a margin of
  // 1e-13 x the sum of the magnitudes of the three terms - 100x their rounding error - so a pair
  // that is dismissed is certainly above T, and ties, NaN and non-positive masses take the exact
  // path.  Counts, minima and partners stay bit-exact.  One branch per source for the thread's
  // targets; inside it every lane loops over its own needs (as OpKick).
  static constexpr int BATCHED = 1;
  template <bool CHECK, int IPT>
  __device__ __forceinline__ static void pairs(const Tgt (&t)[IPT], double (&acc)[IPT][NACC],
                                               const double2 *sj, int jg, const int (&ig)[IPT],
                                               const SweepParams &P) {
    const double2 a = sj[0], b = sj[1], c = sj[2], d = sj[3];
    const double mj = b.y;
    double y[IPT], v2[IPT];
    unsigned need = 0;
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      const double dx = a.x - t[r].x, dy = a.y - t[r].y, dz = b.x - t[r].z;
      double r2 = fma(dx, dx, P.eps2);
      r2 = fma(dy, dy, r2);
      r2 = fma(dz, dz, r2);
      const bool self = CHECK ? (jg == ig[r]) : false;
      if (CHECK) r2 = self ? 1.0 : r2;
      y[r] = rsqrt_nr(r2);
      const double ux = c.x - t[r].vx, uy = c.y - t[r].vy, uz = d.x - t[r].vz;
      v2[r] = fma(uz, uz, fma(uy, uy, ux * ux));
    }
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      const bool self = CHECK ? (jg == ig[r]) : false;
      const double M = t[r].m + mj, mm = t[r].m * mj;
      const double T = fmax(P.dt, acc[r][0]);  // thr or the running minimum, whichever is larger
      const double A = (0.5 * mm) * v2[r], B = mm * (M * y[r]), C = T * M;
      const double lhs = (A - B) - C, mag = (A + B) + fabs(C);
      const bool far = (fma(-1e-13, mag, lhs) > 0.0) & (mj > 0.0) & (t[r].m > 0.0);
      if (!far && !self) need |= 1u << r;
    }
    while (need) {
      const int r = __ffs(need) - 1;
      need &= need - 1;
      Tgt tt = t[0];
#pragma unroll
      for (int rr = 1; rr < IPT; ++rr)
        if (rr == r) tt = t[rr];
      double one[NACC] = {__longlong_as_double(0x7ff0000000000000LL), -1.0, 0.0};
      exact_pair(tt, one, sj, P);
#pragma unroll
      for (int rr = 0; rr < IPT; ++rr)
        if (rr == r) {
          if (one[0] < acc[rr][0] || (one[0] == acc[rr][0] && one[1] < acc[rr][1])) {
            acc[rr][0] = one[0];
            acc[rr][1] = one[1];
          }
          acc[rr][2] += one[2];
        }
    }
  }
  // the exact pair of pair<>() out of line: acc = {e, source index, e < thr ? 1 : 0}
  __device__ __noinline__ static void exact_pair(const Tgt &t, double *one, const double2 *sj,
                                                 const SweepParams &P) {
    one[0] = __longlong_as_double(0x7ff0000000000000LL);
    one[1] = -1.0;
    one[2] = 0.0;
    pair<false>(t, one, sj, nullptr, false, P);
  }
  __device__ __forceinline__ static void combine(double *a, const double *b) {
    if (b[0] < a[0] || (b[0] == a[0] && b[1] >= 0.0 && (a[1] < 0.0 || b[1] < a[1]))) {
      a[0] = b[0];
      a[1] = b[1];
    }
    a[2] += b[2];
  }
  __device__ __forceinline__ static void store(const SweepParams &P, int ig, const double *acc) {
    P.out[0][ig - P.t0] = acc[0];
    P.out[1][ig - P.t0] = acc[1];
    P.out[2][ig - P.t0] = acc[2];
  }
};

// One source against the IPT targets of a thread.  Ops may provide a batched form
// (Op::BATCHED) that walks the IPT interactions phase by phase, so that the instruction
// stream handed to ptxas already interleaves independent interactions and keeps FMAs that
// share a register operand adjacent ("operand bandwidth").
template <class Op, class = void> struct has_batched { static constexpr bool value = false; };
template <class Op> struct has_batched<Op, decltype((void)Op::BATCHED)> {
  static constexpr bool value = true;
};

template <class Op, class = void> struct has_tile_hook { static constexpr bool value = false; };
template <class Op> struct has_tile_hook<Op, decltype((void)Op::TILE_HOOK)> {
  static constexpr bool value = true;
};

template <class Op, class = void> struct naux_of { static constexpr int value = 0; };
template <class Op> struct naux_of<Op, decltype((void)Op::NAUX)> {
  static constexpr int value = Op::NAUX;
};

// sa = this source's slot in the first aux tile; successive aux tiles are TJ*PK doubles apart.
template <class Op, bool CHECK, int IPT>
__device__ __forceinline__ void pairs(const typename Op::Tgt (&tg)[IPT],
                                      double (&acc)[IPT][Op::NACC], const double2 *sj,
                                      const double2 *sa, int jg, const int (&ig)[IPT],
                                      const SweepParams &P) {
  if constexpr (has_batched<Op>::value) {
    Op::template pairs<CHECK, IPT>(tg, acc, sj, jg, ig, P);
  } else {
#pragma unroll
    for (int r = 0; r < IPT; ++r)
      Op::template pair<CHECK>(tg[r], acc[r], sj, sa, jg == ig[r], P);
  }
}

// Batched acc+jerk: same arithmetic as OpGradVDGradV::pair except the last rsqrt step is
// y = y0 + y0 (e p) (two distinct register operands) instead of y0 + p (e y0).
template <int FLAVOUR> struct OpGradVDGradVB : OpGradVDGradV {
  static constexpr int BATCHED = 1;
  template <bool CHECK, int IPT>
  __device__ __forceinline__ static void pairs(const Tgt (&t)[IPT], double (&acc)[IPT][6],
                                               const double2 *sj, int jg, const int (&ig)[IPT],
                                               const SweepParams &P) {
    const double2 a = sj[0], b = sj[1], c = sj[2], d = sj[3];
    double dx[IPT], dy[IPT], dz[IPT], ux[IPT], uy[IPT], uz[IPT], r2[IPT], rv[IPT], mj[IPT];
    double y[IPT], w[IPT], a3[IPT];
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      dx[r] = a.x - t[r].x;
      dy[r] = a.y - t[r].y;
      dz[r] = b.x - t[r].z;
    }
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      r2[r] = fma(dx[r], dx[r], P.eps2);
      r2[r] = fma(dy[r], dy[r], r2[r]);
      r2[r] = fma(dz[r], dz[r], r2[r]);
      mj[r] = b.y;
      if (CHECK) {
        const bool self = jg == ig[r];
        r2[r] = self ? 1.0 : r2[r];
        mj[r] = self ? 0.0 : mj[r];
      }
    }
    double y0[IPT];
#pragma unroll
    for (int r = 0; r < IPT; ++r) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[r]) : "d"(r2[r]));
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      ux[r] = c.x - t[r].vx;
      uy[r] = c.y - t[r].vy;
      uz[r] = d.x - t[r].vz;
    }
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      rv[r] = dx[r] * ux[r];
      rv[r] = fma(dy[r], uy[r], rv[r]);
      rv[r] = fma(dz[r], uz[r], rv[r]);
    }
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      double tt = r2[r] * y0[r];
      double e = fma(-tt, y0[r], 1.0);
      double pp = fma(0.375, e, 0.5);
      if (FLAVOUR & 1) {
        double q = e * pp;
        y[r] = fma(y0[r], q, y0[r]);
      } else {
        double q = e * y0[r];
        y[r] = fma(pp, q, y0[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      double y2 = y[r] * y[r];
      w[r] = (mj[r] * y[r]) * y2;
      a3[r] = (rv[r] * y2) * -3.0;
    }
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      acc[r][0] = fma(w[r], dx[r], acc[r][0]);
      acc[r][1] = fma(w[r], dy[r], acc[r][1]);
      acc[r][2] = fma(w[r], dz[r], acc[r][2]);
    }
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      ux[r] = fma(a3[r], dx[r], ux[r]);
      uy[r] = fma(a3[r], dy[r], uy[r]);
      uz[r] = fma(a3[r], dz[r], uz[r]);
    }
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      acc[r][3] = fma(w[r], ux[r], acc[r][3]);
      acc[r][4] = fma(w[r], uy[r], acc[r][4]);
      acc[r][5] = fma(w[r], uz[r], acc[r][5]);
    }
  }
};

// Tangent (first-order forward mode) of the acc / acc+jerk sums for KT directions per pass:
// DFloatBase.grad_v / grad_v_dgrad_v (dFloat_advancer.ml:57-81) summed as in
// dFloat_hermite.ml:177-186.  Direction k's packed array aux[k] holds per body
// {dq_x, dq_y, dq_z, 0, dv_x, dv_y, dv_z, 0}, dv = dp / m (masses and eps2 carry no tangent,
// dFloat_hermite.ml:36-46).  With e = dq_j - dq_i, g = dv_j - dv_i, c1 = d.e, c2 = e.u + d.g:
//   dA = sum w (e - 3 y^2 c1 d)
//   dJ = sum w (g + a3 e + da3 d - 3 y^2 c1 (u + a3 d)),  da3 = -3 y^2 c2 + 6 (d.u) y^4 c1
// (checked against central differences in tests/).  Accumulators hold the
// tangents only; the value sums come from the plain kernels.
template <int KT, bool JERK> struct OpTangent {
  static constexpr int NAUX = KT;
  static constexpr int NC = JERK ? 6 : 3;
  static constexpr int NACC = NC * KT;
  struct Tgt {
    double x, y, z, vx, vy, vz;
    double e[KT][3], g[KT][3];
  };
  __device__ __forceinline__ static void load(Tgt &t, const double *pk, const SweepParams &P,
                                              int i) {
    t.x = pk[0];
    t.y = pk[1];
    t.z = pk[2];
    t.vx = pk[4];
    t.vy = pk[5];
    t.vz = pk[6];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const double *a = P.aux + (size_t)k * P.aux_stride + (size_t)i * PK;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        t.e[k][c] = a[c];
        t.g[k][c] = JERK ? a[4 + c] : 0.0;
      }
    }
  }
  __device__ __forceinline__ static void zero(double *acc) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  }
  template <bool CHECK>
  __device__ __forceinline__ static void pair(const Tgt &t, double *acc, const double2 *sj,
                                              const double2 *sa, bool self,
                                              const SweepParams &P) {
    const double2 a = sj[0], b = sj[1];
    const double d[3] = {a.x - t.x, a.y - t.y, b.x - t.z};
    double r2 = fma(d[0], d[0], P.eps2);
    r2 = fma(d[1], d[1], r2);
    r2 = fma(d[2], d[2], r2);
    double mj = b.y;
    if (CHECK) {
      r2 = self ? 1.0 : r2;
      mj = self ? 0.0 : mj;
    }
    const double y = rsqrt_nr(r2);
    const double y2 = y * y;
    const double w = (mj * y) * y2;
    const double m3y2 = -3.0 * y2;
    double u[3] = {0.0, 0.0, 0.0}, tt[3] = {0.0, 0.0, 0.0}, rv = 0.0, a3 = 0.0;
    if (JERK) {
      const double2 c = sj[2], dd = sj[3];
      u[0] = c.x - t.vx;
      u[1] = c.y - t.vy;
      u[2] = dd.x - t.vz;
      rv = fma(d[2], u[2], fma(d[1], u[1], d[0] * u[0]));
      a3 = rv * m3y2;
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3) tt[c3] = fma(a3, d[c3], u[c3]);
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const double2 *ak = sa + (size_t)k * (TJ * PK / 2);
      const double2 a0 = ak[0], a1 = ak[1];
      const double e[3] = {a0.x - t.e[k][0], a0.y - t.e[k][1], a1.x - t.e[k][2]};
      const double c1 = fma(d[2], e[2], fma(d[1], e[1], d[0] * e[0]));
      const double b1 = m3y2 * c1;
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3)
        acc[k * NC + c3] = fma(w, fma(b1, d[c3], e[c3]), acc[k * NC + c3]);
      if (JERK) {
        const double2 a2 = ak[2], a3v = ak[3];
        const double g[3] = {a2.x - t.g[k][0], a2.y - t.g[k][1], a3v.x - t.g[k][2]};
        double c2 = fma(e[2], u[2], fma(e[1], u[1], e[0] * u[0]));
        c2 = fma(d[2], g[2], fma(d[1], g[1], fma(d[0], g[0], c2)));
        // da3 = -3 y2 (c2 - 2 rv y2 c1)
        const double da3 = m3y2 * fma(-2.0 * (rv * y2), c1, c2);
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
          double s = fma(b1, tt[c3], g[c3]);
          s = fma(a3, e[c3], s);
          s = fma(da3, d[c3], s);
          acc[k * NC + 3 + c3] = fma(w, s, acc[k * NC + 3 + c3]);
        }
      }
    }
  }
  __device__ __forceinline__ static void combine(double *a, const double *b) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) a[k] += b[k];
  }
  // out[0] = gsum_tan [K][3 nt], out[1] = dgsum_tan [K][3 nt], already offset to this pass.
  __device__ __forceinline__ static void store(const SweepParams &P, int ig, const double *acc) {
    const double s = -P.packed[(size_t)ig * PK + 3];
    const size_t nt3 = 3 * (size_t)(P.t1 - P.t0);
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      double *g = P.out[0] + (size_t)k * nt3 + 3 * (size_t)(ig - P.t0);
#pragma unroll
      for (int c = 0; c < 3; ++c) g[c] = s * acc[k * NC + c];
      if (JERK) {
        double *dg = P.out[1] + (size_t)k * nt3 + 3 * (size_t)(ig - P.t0);
#pragma unroll
        for (int c = 0; c < 3; ++c) dg[c] = s * acc[k * NC + 3 + c];
      }
    }
  }
};

// ---- the sweep kernel ----------------------------------------------------------------------
constexpr size_t sweep_smem_bytes(int naux) {
  return (size_t)STAGES * TJ * PK * sizeof(double) * (1 + naux) + 2 * STAGES * 8;
}

// NT threads per CTA, IPT targets per thread (i-tile TI = NT*IPT), MINB resident CTAs per SM
// asked of the register allocator, UNROLL = source-loop unroll of the off-diagonal path.
// PEER: source tiles are bulk-copied straight out of the owning GPU's memory over NVLink
// (SweepParams::shard) - the exchange step of the target-sharded multi-GPU sweep fused into the
// sweep itself, tile by tile behind the same TMA ring, instead of an all-gather in front of it.
template <class Op, int NT, int IPT, int MINB, int UNROLL, bool PEER = false>
__global__ void __launch_bounds__(NT, MINB) sweep_kernel(const SweepParams P) {
  constexpr int TI = NT * IPT;
  constexpr int NAUX = naux_of<Op>::value;
  constexpr int STAGE_DOUBLES = TJ * PK * (1 + NAUX);  // base tile, then NAUX aux tiles
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tiles = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * STAGE_DOUBLES * 8);
  uint64_t *empty = full + STAGES;

  const int tid = threadIdx.x;
  const long long u_begin = (long long)blockIdx.x * P.units_per_cta;
  long long u_end = u_begin + P.units_per_cta;
  if (u_end > P.units) u_end = P.units;
  const int ntiles = (int)(u_end - u_begin);  // source tiles this CTA streams (>= 1)

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NT / 32);
    }
    mbar_fence_init();
    if constexpr (PEER) {
      // Every owner's rows must be published before this CTA pulls its first tile.  Stream-K
      // hands a CTA tiles of every owner from the start, so waiting owner by owner inside the
      // loop would buy nothing - and keeping that state alive across the hot loop costs it
      // registers (measured: +0.8 % kernel time).
      if (P.ready != nullptr) {
        const int r_lo = P.s0 / P.shard_rows, r_hi = (P.s1 - 1) / P.shard_rows;
        // A peer that never publishes (it crashed) must not hang this GPU: after P.wait_ns
        // (nbk_set_peer_timeout, default 20 s; 0 = wait forever, e.g. under a debugger) the
        // kernel traps, which the host sees as a launch failure on its next CUDA call.  All
        // ranks must call their sweeps in lockstep (same number of epochs).
        unsigned long long t_start;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
        for (int r = r_lo; r <= r_hi; ++r)  // the owners of sources [s0, s1)
          if (r != P.self_rank)
            while (ld_acquire_sys(P.ready + r) < P.epoch) {
              __nanosleep(64);
              unsigned long long t_now;
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
              if (P.wait_ns != 0 && t_now - t_start > P.wait_ns) __trap();
            }
        fence_proxy_async();
      }
    }
  }
  __syncthreads();

  // Producer side (thread 0): tile k of this CTA is unit u_begin+k -> source tile jt.
  auto issue = [&](int k) {
    long long u = u_begin + k;
    int jt = (int)(u % P.n_jt);
    int j0 = P.s0 + jt * TJ;
    int cnt = min(TJ, P.s1 - j0);
    int st = k % STAGES;
    uint32_t bytes = (uint32_t)cnt * PK * 8;
    double *dst = tiles + (size_t)st * STAGE_DOUBLES;
    const double *src = P.packed + (size_t)j0 * PK;
    const double *asrc = P.aux + (size_t)j0 * PK;
    if constexpr (PEER) {
      // shard_rows and s0 are multiples of TJ: a tile never straddles two owners
      const int owner = j0 / P.shard_rows;
      const size_t row = (size_t)(j0 - owner * P.shard_rows);
      src = P.shard[owner] + row * PK;
      asrc = P.aux_shard[owner] + row * PK;
    }
    mbar_arrive_expect_tx(&full[st], bytes * (1 + NAUX));
    tma_bulk_g2s(dst, src, bytes, &full[st]);
#pragma unroll
    for (int a = 0; a < NAUX; ++a)
      tma_bulk_g2s(dst + (size_t)(1 + a) * TJ * PK, asrc + (size_t)a * P.aux_stride, bytes,
                   &full[st]);
  };
  if (tid == 0) {
    for (int k = 0; k < STAGES - 1 && k < ntiles; ++k) issue(k);
  }

  typename Op::Tgt tg[IPT];
  double acc[IPT][Op::NACC];
  int ig[IPT];
  int seg_it = -1, seg_jt0 = 0;
  const int head_it = (int)(u_begin / P.n_jt);

  auto flush = [&](int it, int jt0, int jt_last) {
    const bool complete = (jt0 == 0) && (jt_last == P.n_jt - 1);
    if (complete) {
#pragma unroll
      for (int r = 0; r < IPT; ++r) {
        int i = P.t0 + it * TI + r * NT + tid;
        if (i < P.t1) Op::store(P, i, acc[r]);
      }
    } else {
      const int slot = 2 * blockIdx.x + (it == head_it ? 0 : 1);
      double *dst = P.partial + (size_t)slot * Op::NACC * TI;
#pragma unroll
      for (int r = 0; r < IPT; ++r)
#pragma unroll
        for (int c = 0; c < Op::NACC; ++c) dst[c * TI + r * NT + tid] = acc[r][c];
    }
  };

  for (int k = 0; k < ntiles; ++k) {
    const long long u = u_begin + k;
    const int it = (int)(u / P.n_jt);
    const int jt = (int)(u - (long long)it * P.n_jt);
    if (it != seg_it) {
      if (seg_it >= 0) flush(seg_it, seg_jt0, P.n_jt - 1);
      seg_it = it;
      seg_jt0 = jt;
#pragma unroll
      for (int r = 0; r < IPT; ++r) {
        int i = P.t0 + it * TI + r * NT + tid;
        ig[r] = i;
        int ic = i < P.t1 ? i : P.t1 - 1;  // clamp: lanes past the end compute, never store
        Op::load(tg[r], P.packed + (size_t)ic * PK, P, ic);
        Op::zero(acc[r]);
      }
      // a run that starts late adopts what the earlier runs of its targets already know
      if constexpr (has_tile_hook<Op>::value) Op::template tile_end<IPT>(acc, ig, P);
    }
    // keep STAGES-1 tiles in flight: refill the stage tile k-1 just vacated
    if (tid == 0) {
      int kk = k + STAGES - 1;
      if (kk < ntiles) {
        if (kk >= STAGES) mbar_wait(&empty[kk % STAGES], ((kk / STAGES) & 1) ^ 1);
        issue(kk);
      }
    }
    const int st = k % STAGES;
    mbar_wait(&full[st], (k / STAGES) & 1);

    const int j0 = P.s0 + jt * TJ;
    const int cnt = min(TJ, P.s1 - j0);
    const double2 *sj = reinterpret_cast<const double2 *>(tiles + (size_t)st * STAGE_DOUBLES);
    const double2 *sa = sj + TJ * (PK / 2);
    const int i_lo = P.t0 + it * TI;
    const bool diag = (j0 < i_lo + TI) && (j0 + cnt > i_lo);
    if (diag) {
#pragma unroll 2
      for (int jj = 0; jj < cnt; ++jj)
        pairs<Op, true, IPT>(tg, acc, sj + jj * (PK / 2), sa + jj * (PK / 2), j0 + jj, ig, P);
    } else if (has_tile_hook<Op>::value && jt == seg_jt0) {
      // the first tile of a run: exchange every 16 sources (see Op::tile_end)
#pragma unroll 1
      for (int jj = 0; jj < cnt; ++jj) {
        pairs<Op, false, IPT>(tg, acc, sj + jj * (PK / 2), sa + jj * (PK / 2), j0 + jj, ig, P);
        if constexpr (has_tile_hook<Op>::value) {
          if ((jj & 15) == 15) Op::template tile_end<IPT>(acc, ig, P);
        }
      }
    } else {
#pragma unroll UNROLL
      for (int jj = 0; jj < cnt; ++jj)
        pairs<Op, false, IPT>(tg, acc, sj + jj * (PK / 2), sa + jj * (PK / 2), j0 + jj, ig, P);
    }
    // exchange with the other runs of these targets: at every tile while a run is young (that is
    // when its own bound is still loose), then every eighth tile (the loads cost a stall)
    if constexpr (has_tile_hook<Op>::value) {
      const int kseg = jt - seg_jt0;
      if (kseg < 4 || (kseg & 7) == 7) Op::template tile_end<IPT>(acc, ig, P);
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[st]);
  }
  // last run: ends at the CTA's last unit
  {
    const long long ul = u_end - 1;
    const int jt_last = (int)(ul - (long long)seg_it * P.n_jt);
    flush(seg_it, seg_jt0, jt_last);
  }
}

// Fix-up: combine, per i-tile, the partial slots of the CTAs that shared its source range,
// in ascending CTA (= ascending source) order.  One thread per target.
template <class Op, int TI>
__global__ void __launch_bounds__(TI) fixup_kernel(const SweepParams P) {
  const int it = blockIdx.x;
  const long long u_lo = (long long)it * P.n_jt, u_hi = u_lo + P.n_jt - 1;
  const int c_lo = (int)(u_lo / P.units_per_cta), c_hi = (int)(u_hi / P.units_per_cta);
  if (c_lo == c_hi) return;  // the tile ran complete inside one CTA: already stored
  const int lane = threadIdx.x;
  const int i = P.t0 + it * TI + lane;
  if (i >= P.t1) return;
  double acc[Op::NACC], part[Op::NACC];
  Op::zero(acc);
  for (int c = c_lo; c <= c_hi; ++c) {
    const int head_it = (int)(((long long)c * P.units_per_cta) / P.n_jt);
    const int slot = 2 * c + (it == head_it ? 0 : 1);
    const double *src = P.partial + (size_t)slot * Op::NACC * TI;
#pragma unroll
    for (int k = 0; k < Op::NACC; ++k) part[k] = src[k * TI + lane];
    Op::combine(acc, part);
  }
  Op::store(P, i, acc);
}

// Publishes "rank `self`'s rows of epoch `epoch` are in place" into every rank's ready array
// (peer-mapped): launched on the stream after the kernel that wrote the rows.
struct PublishParams {
  unsigned long long *ready_of[MAX_PEERS];
};
__global__ void publish_kernel(PublishParams F, int nranks, int self, unsigned long long epoch) {
  const int r = threadIdx.x;
  if (r < nranks) {
    __threadfence_system();
    st_release_sys(F.ready_of[r] + self, epoch);
  }
}

// ---- few targets: source-parallel kernel -----------------------------------------------------
// Hermite.advance_body (one target, hermite.ml:106-115) and the small active blocks of
// Advancer.A / Leapfrog have 1..a few dozen targets: too few for i-parallel lanes.  Here every
// thread owns SOURCES (grid-stride over [s0,s1), one coalesced 64-byte read each), loops over a
// group of NTG targets held in registers, and the per-target sums are reduced across the warp
// with shuffles, across warps through shared memory, and across CTAs by small_fix_kernel - all
// in a fixed order.  partial layout: [target][cta][NACC].
template <class Op, int NTG>
__global__ void __launch_bounds__(256) small_kernel(const SweepParams P, int tg0, int ncta) {
  static_assert(naux_of<Op>::value == 0, "source-parallel kernel has no aux tiles");
  __shared__ double red[8][NTG * Op::NACC];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  typename Op::Tgt tg[NTG];
  double acc[NTG][Op::NACC];
  int ig[NTG];
#pragma unroll
  for (int r = 0; r < NTG; ++r) {
    int i = tg0 + r;
    ig[r] = i < P.t1 ? i : -1;
    int ic = i < P.t1 ? i : P.t1 - 1;
    Op::load(tg[r], P.packed + (size_t)ic * PK, P, ic);
    Op::zero(acc[r]);
  }
  for (int j = P.s0 + blockIdx.x * 256 + tid; j < P.s1; j += ncta * 256) {
    double2 src[PK / 2];
    const double2 *gp = reinterpret_cast<const double2 *>(P.packed + (size_t)j * PK);
#pragma unroll
    for (int c = 0; c < PK / 2; ++c) src[c] = gp[c];
#pragma unroll
    for (int r = 0; r < NTG; ++r) Op::template pair<true>(tg[r], acc[r], src, src, j == ig[r], P);
  }
#pragma unroll
  for (int r = 0; r < NTG; ++r) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      double other[Op::NACC];
#pragma unroll
      for (int c = 0; c < Op::NACC; ++c) other[c] = __shfl_down_sync(0xffffffffu, acc[r][c], off);
      Op::combine(acc[r], other);
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < Op::NACC; ++c) red[warp][r * Op::NACC + c] = acc[r][c];
    }
  }
  __syncthreads();
  if (tid < NTG && tg0 + tid < P.t1) {
    double a[Op::NACC], b[Op::NACC];
#pragma unroll
    for (int c = 0; c < Op::NACC; ++c) a[c] = red[0][tid * Op::NACC + c];
    for (int w = 1; w < 8; ++w) {
#pragma unroll
      for (int c = 0; c < Op::NACC; ++c) b[c] = red[w][tid * Op::NACC + c];
      Op::combine(a, b);
    }
    double *dst = P.partial + ((size_t)(tg0 + tid - P.t0) * ncta + blockIdx.x) * Op::NACC;
#pragma unroll
    for (int c = 0; c < Op::NACC; ++c) dst[c] = a[c];
  }
}

template <class Op> __global__ void small_fix_kernel(const SweepParams P, int ncta) {
  const int i = P.t0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.t1) return;
  double a[Op::NACC], b[Op::NACC];
  Op::zero(a);
  const double *src = P.partial + (size_t)(i - P.t0) * ncta * Op::NACC;
  for (int c = 0; c < ncta; ++c) {
#pragma unroll
    for (int k = 0; k < Op::NACC; ++k) b[k] = src[(size_t)c * Op::NACC + k];
    Op::combine(a, b);
  }
  Op::store(P, i, a);
}

// ---- Hermite.advance_body on device-resident state ------------------------------------------
// Resident Hermite.b (hermite.ml:20-29) as 16 doubles per body:
// {t, m, q[3], p[3], f[3], df[3], 0, 0}.  One launch = one force evaluation of
// Hermite.advance_body (hermite.ml:95-115): optionally overwrite body `upd` with the result
// of the previous step (passed by value, so no separate copy or launch), predict every body
// to time nt with the reference's operation sequence (hermite.ml:53-72), sum
// grad_v_dgrad_v(target i, j) over j != i source-parallel, reduce in fixed order; the last
// CTA to finish (ticket counter) adds the per-CTA partials in CTA order and writes the result.
constexpr int HB = 16;
struct HermiteUpd {
  double s[13];  // t, q[3], p[3], f[3], df[3]
};

__device__ __forceinline__ void hermite_predict(const double *r, double nt, double &m,
                                                double (&qp)[3], double (&vp)[3]) {
  m = r[1];
  const double dt = __dsub_rn(nt, r[0]);
  const double pc = __ddiv_rn(dt, m);
  const double fc = __dmul_rn(__dmul_rn(pc, dt), 0.5);
  const double dfc = __ddiv_rn(__dmul_rn(fc, dt), 3.0);
  const double pdfc = __dmul_rn(__dmul_rn(dt, dt), 0.5);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double qk = r[2 + k], pk = r[5 + k], fk = r[8 + k], dfk = r[11 + k];
    qp[k] = __dadd_rn(__dadd_rn(__dadd_rn(qk, __dmul_rn(pc, pk)), __dmul_rn(fc, fk)),
                      __dmul_rn(dfc, dfk));
    const double pp = __dadd_rn(__dadd_rn(pk, __dmul_rn(dt, fk)), __dmul_rn(pdfc, dfk));
    vp[k] = __ddiv_rn(pp, m);
  }
}

// Same predictor with 1/m and 1/3 as multiplications (no IEEE divisions): a few ulp from the
// reference's sequence, used only inside the device-resident stepping loop where the five
// divisions per body per step are the critical path.
__device__ __forceinline__ void hermite_predict_fast(const double *r, double nt, double &m,
                                                     double (&qp)[3], double (&vp)[3]) {
  m = r[1];
  const double im = rcp_nr(m);
  const double dt = nt - r[0];
  const double pc = dt * im;
  const double fc = (pc * dt) * 0.5;
  const double dfc = (fc * dt) * (1.0 / 3.0);
  const double pdfc = (dt * dt) * 0.5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double qk = r[2 + k], pk = r[5 + k], fk = r[8 + k], dfk = r[11 + k];
    qp[k] = fma(dfc, dfk, fma(fc, fk, fma(pc, pk, qk)));
    vp[k] = fma(pdfc, dfk, fma(dt, fk, pk)) * im;
  }
}

__global__ void __launch_bounds__(256) hermite_body_kernel(double *__restrict__ state, int n,
                                                           int it, double nt, double eps2,
                                                           int upd, const HermiteUpd u,
                                                           double *__restrict__ partial,
                                                           unsigned int *__restrict__ ticket,
                                                           double *out, double seq) {
  __shared__ double red[8][6];
  __shared__ bool last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.x * 256 + tid;
  SweepParams P;
  P.eps2 = eps2;
  auto fetch = [&](int b, double *r) {
    const double2 *g = reinterpret_cast<const double2 *>(state + (size_t)b * HB);
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      double2 v = g[c];
      r[2 * c] = v.x;
      r[2 * c + 1] = v.y;
    }
    if (b == upd) {
      r[0] = u.s[0];
#pragma unroll
      for (int c = 0; c < 12; ++c) r[2 + c] = u.s[1 + c];
    }
  };
  double acc[6] = {0, 0, 0, 0, 0, 0};
  {
    double rt[14], mt, qt[3], vt[3];
    fetch(it, rt);
    hermite_predict(rt, nt, mt, qt, vt);
    OpGradVDGradV::Tgt tg = {qt[0], qt[1], qt[2], vt[0], vt[1], vt[2]};
    if (j < n) {
      double rj[14], mj, qj[3], vj[3];
      fetch(j, rj);
      if (j == upd) {  // persist the previous step's result for later launches
        double *w = state + (size_t)j * HB;
        w[0] = rj[0];
#pragma unroll
        for (int c = 0; c < 12; ++c) w[2 + c] = rj[2 + c];
      }
      hermite_predict(rj, nt, mj, qj, vj);
      const double2 src[4] = {make_double2(qj[0], qj[1]), make_double2(qj[2], mj),
                              make_double2(vj[0], vj[1]), make_double2(vj[2], 0.0)};
      OpGradVDGradV::pair<true>(tg, acc, src, src, j == it, P);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], off);
  if (lane == 0)
#pragma unroll
    for (int c = 0; c < 6; ++c) red[warp][c] = acc[c];
  __syncthreads();
  if (tid < 6) {
    double a = red[0][tid];
    for (int w = 1; w < 8; ++w) a += red[w][tid];
    partial[(size_t)blockIdx.x * 6 + tid] = a;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && tid < 6) {
    __threadfence();
    double a = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) a += __ldcg(&partial[(size_t)b * 6 + tid]);
    // gsum = -m_i acc (gradient convention of Base.grad_v_dgrad_v)
    const double mi = state[(size_t)it * HB + 1];
    // `out` is host memory mapped into the device (zero-copy): the host spins on out[7] == seq
    // instead of paying a D2H copy and a stream synchronisation per step.
    out[tid] = -mi * a;
    __threadfence_system();
    if (tid == 0) *ticket = 0u;
  }
  if (last) {
    __syncwarp();
    if (tid == 0) {
      __threadfence_system();
      *reinterpret_cast<volatile double *>(out + 7) = seq;
    }
  }
}

// ---- Hermite.advance_bodies + finish_bs entirely on the device ---------------------------------
// The reference steps ONE body at a time (hermite.ml:171-178): pop the body with the earliest
// next time, predict everybody to that time, 1 x (N-1) acc+jerk, correct, new time step,
// re-insert.  A kernel launch per step costs ~10 us; this persistent single-CTA kernel keeps
// the whole Hermite.b array in SHARED MEMORY (136 B per body, N <= HRES_MAX_N) and runs the loop
// on the device: per step two block reductions (arg-min of next time, 6 force sums), one
// thread doing the corrector - 3.6 us per step at N = 1024 (hermite_resident_spec_kernel below is
// the pipelined version, 2.0 us).
// Scheduling follows the reference's list semantics: earliest next time first, and among equal
// next times the most recently re-inserted first (List.merge puts the new body before equal
// elements, hermite.ml:178), initial ties by array order (stable sort, :187); finish_bs
// (hermite.ml:132-143) then steps every body to stop_time in that same order.
// Record slots: 0 t, 1 m, 2-4 q, 5-7 p, 8-10 f, 11-13 df, 14 h, 15 re-insertion stamp.
// The scheduler's arg-min as integer reductions.  A body's rank in the reference's list is
// (next time ascending, re-insertion stamp DESCENDING, index ascending); with the order-preserving
// map of a double onto an unsigned integer below that is the lexicographic minimum of two 64-bit
// keys, and a 64-bit minimum over a warp is two redux.sync (high word, then the low word among the
// lanes that hold the minimal high word) instead of five shuffle-compare-select rounds on three
// values.  Measured in hermite_resident_spec_kernel: 3900 -> ~500 cycles per selection.
__device__ __forceinline__ unsigned long long hkey_of_double(double x) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double hkey_to_double(unsigned long long k) {
  return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
constexpr unsigned long long HKEY_NONE = ~0ull;
constexpr int HKEY_IDX_BITS = 20;
__device__ __forceinline__ unsigned long long hkey_tie(double stamp, int idx) {
  // stamps are step counts (< 2^43) kept as doubles; newer (larger) first, then the lower index
  return ((0x7ffffffffffull - (unsigned long long)stamp) << HKEY_IDX_BITS) | (unsigned)idx;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
  const unsigned hi = (unsigned)(v >> 32);
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned lo = (hi == mh) ? (unsigned)v : 0xffffffffu;
  const unsigned ml = __reduce_min_sync(0xffffffffu, lo);
  return ((unsigned long long)mh << 32) | ml;
}
// lexicographic minimum of (k1, k2) over the warp
__device__ __forceinline__ void warp_min_pair(unsigned long long &k1, unsigned long long &k2) {
  const unsigned long long m1 = warp_min_u64(k1);
  k2 = warp_min_u64(k1 == m1 ? k2 : HKEY_NONE);
  k1 = m1;
}

constexpr int HRES_STRIDE = 17;   // odd stride: conflict-free per-thread rows
constexpr int HRES_MAX_N = 1536;

template <int HRES_THREADS>
__global__ void __launch_bounds__(HRES_THREADS, 1)
    hermite_resident_kernel(double *__restrict__ state, int n, double stop_time, double eta,
                            double eps2, long long max_steps, long long *__restrict__ nsteps_out) {
  extern __shared__ double rec[];  // [n][HRES_STRIDE]; slot 16 = done flag
  __shared__ double red[HRES_THREADS / 32][6];
  __shared__ double key_nt[HRES_THREADS / 32], key_st[HRES_THREADS / 32];
  __shared__ int key_idx[HRES_THREADS / 32];
  __shared__ int sel_idx;
  __shared__ double sel_nt;
  constexpr int NW = HRES_THREADS / 32;
  constexpr int NONE = 0x7fffffff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  SweepParams P;
  P.eps2 = eps2;
  for (int j = tid; j < n; j += HRES_THREADS) {
#pragma unroll
    for (int c = 0; c < 16; ++c) rec[j * HRES_STRIDE + c] = state[(size_t)j * HB + c];
    rec[j * HRES_STRIDE + 16] = 0.0;
  }
  __syncthreads();
  long long steps = 0, stamp = 0;
  bool finishing = false;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  auto better = [](double ont, double ost, int oidx, double knt, double kst, int kidx) {
    return oidx != NONE &&
           (kidx == NONE || (ont < knt) ||
            (ont == knt && (ost > kst || (ost == kst && oidx < kidx))));
  };
  for (;;) {
    // ---- 1. arg-min over unfinished bodies of (next time asc, stamp desc, index asc)
    double knt = INF, kst = -INF;
    int kidx = NONE;
    for (int j = tid; j < n; j += HRES_THREADS) {
      const double *r = rec + j * HRES_STRIDE;
      if (r[16] != 0.0) continue;
      const double nt = r[0] + r[14];
      if (better(nt, r[15], j, knt, kst, kidx)) {
        knt = nt;
        kst = r[15];
        kidx = j;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ont = __shfl_down_sync(0xffffffffu, knt, off);
      const double ost = __shfl_down_sync(0xffffffffu, kst, off);
      const int oidx = __shfl_down_sync(0xffffffffu, kidx, off);
      if (better(ont, ost, oidx, knt, kst, kidx)) {
        knt = ont;
        kst = ost;
        kidx = oidx;
      }
    }
    if (lane == 0) {
      key_nt[warp] = knt;
      key_st[warp] = kst;
      key_idx[warp] = kidx;
    }
    __syncthreads();
    if (warp == 0) {
      knt = lane < NW ? key_nt[lane] : INF;
      kst = lane < NW ? key_st[lane] : -INF;
      kidx = lane < NW ? key_idx[lane] : NONE;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ont = __shfl_down_sync(0xffffffffu, knt, off);
        const double ost = __shfl_down_sync(0xffffffffu, kst, off);
        const int oidx = __shfl_down_sync(0xffffffffu, kidx, off);
        if (better(ont, ost, oidx, knt, kst, kidx)) {
          knt = ont;
          kst = ost;
          kidx = oidx;
        }
      }
      if (lane == 0) {
        sel_idx = kidx;
        sel_nt = knt;
      }
    }
    __syncthreads();
    const int it = sel_idx;
    if (it == NONE) break;  // every body finished
    if (!finishing) {
      if (steps >= max_steps) break;
      if (sel_nt > stop_time) finishing = true;  // advance_bodies -> finish_bs (:172-173)
    }
    double *tr = rec + it * HRES_STRIDE;
    // finish_bs: {b with h = stop_time - b.t}; every thread computes the same value
    const double h = finishing ? __dsub_rn(stop_time, tr[0]) : tr[14];
    const double nt = __dadd_rn(tr[0], h);
    double mt, qt[3], vt[3];
    hermite_predict_fast(tr, nt, mt, qt, vt);
    // ---- 2. 1 x (N-1) acc+jerk over predicted bodies (hermite.ml:106-115)
    const OpGradVDGradV::Tgt tg = {qt[0], qt[1], qt[2], vt[0], vt[1], vt[2]};
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int j = tid; j < n; j += HRES_THREADS) {
      double mj, qj[3], vj[3];
      hermite_predict_fast(rec + j * HRES_STRIDE, nt, mj, qj, vj);
      const double2 src[4] = {make_double2(qj[0], qj[1]), make_double2(qj[2], mj),
                              make_double2(vj[0], vj[1]), make_double2(vj[2], 0.0)};
      OpGradVDGradV::pair<true>(tg, acc, src, src, j == it, P);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int c = 0; c < 6; ++c) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], off);
    if (lane == 0)
#pragma unroll
      for (int c = 0; c < 6; ++c) red[warp][c] = acc[c];
    __syncthreads();
    // ---- 3. one thread: corrector (hermite.ml:116-123), new time step (hermite.ml:87-93)
    ++steps;
    ++stamp;
    if (tid == 0) {
      double fp[3], dfp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double a = 0.0, d = 0.0;
        for (int w = 0; w < NW; ++w) {
          a += red[w][c];
          d += red[w][3 + c];
        }
        fp[c] = mt * a;  // f = -gsum = m_i * acc
        dfp[c] = mt * d;
      }
      const double m = mt;
      double qn[3], pn[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double q0 = tr[2 + c], p0 = tr[5 + c], f0 = tr[8 + c], df0 = tr[11 + c];
        pn[c] = __dadd_rn(__dadd_rn(p0, __dmul_rn(__dmul_rn(0.5, h), __dadd_rn(f0, fp[c]))),
                          __dmul_rn(__ddiv_rn(__dmul_rn(h, h), 12.0), __dsub_rn(df0, dfp[c])));
        qn[c] = __dadd_rn(
            __dadd_rn(q0, __dmul_rn(__ddiv_rn(__dmul_rn(0.5, h), m), __dadd_rn(p0, pn[c]))),
            __dmul_rn(__ddiv_rn(__dmul_rn(h, h), __dmul_rn(12.0, m)), __dsub_rn(f0, fp[c])));
      }
      // timescale eta m q_new qp fp h, with qp = the target's predicted position
      const double dx = qn[0] - qt[0], dy = qn[1] - qt[1], dz = qn[2] - qt[2];
      const double rr = __dsqrt_rn(
          __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
      const double fn = __dsqrt_rn(__dadd_rn(
          __dadd_rn(__dmul_rn(fp[0], fp[0]), __dmul_rn(fp[1], fp[1])), __dmul_rn(fp[2], fp[2])));
      double hn;
      if (rr == 0.0) {
        hn = __dmul_rn(h, 1.189207115002721);  // 2 ** 0.25
      } else {
        const double x =
            __ddiv_rn(__dmul_rn(__dmul_rn(eta, __dmul_rn(h, h)), fn), __dmul_rn(m, rr));
        hn = __dmul_rn(h, __dsqrt_rn(__dsqrt_rn(x)));  // x ** 0.25
      }
      tr[0] = nt;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        tr[2 + c] = qn[c];
        tr[5 + c] = pn[c];
        tr[8 + c] = fp[c];
        tr[11 + c] = dfp[c];
      }
      tr[14] = hn;
      tr[15] = (double)stamp;
      if (finishing) tr[16] = 1.0;
    }
    __syncthreads();
  }
  __syncthreads();
  for (int j = tid; j < n; j += HRES_THREADS)
#pragma unroll
    for (int c = 0; c < 16; ++c) state[(size_t)j * HB + c] = rec[j * HRES_STRIDE + c];
  if (tid == 0) *nsteps_out = steps;
}

// ---- the same loop, software-pipelined -------------------------------------------------------------
// In hermite_resident_kernel everything serialises behind one thread: select, sum, correct,
// select...  Here three things run side by side in every step k, on bodies whose records are not
// being written:
//   thread 0     the corrector and new time step of body i_k (a chain of dependent FP64 operations
//                with three divisions and four square roots, ~1800 cycles)
//   other warps  the force sum of step k+1: its target r (the earliest unfinished body other than
//                i_k) is known when the step starts - see below - so everybody else is predicted
//                to r's time and the 1 x (N-2) interactions that do not involve i_k are summed
//   one warp     the two earliest unfinished bodies other than i_k, b1 (= r) and b2
// After the step's single barrier the schedule's rule decides who is next: i_k again only if its
// new next time is <= r's (ties go to the most recently re-inserted body, hermite.ml:178) - then
// the speculative sums are dropped and the step is done the plain way (10 of 181 843 steps at
// N = 1024); otherwise r, for which thread 0 only has to add the one missing interaction (r
// against the NEW i_k) before its corrector.  The target after next needs no search either:
// excluding r, the earliest body is b2 or the re-inserted i_k, whichever comes first.
// Keys: a body's rank (next time asc, re-insertion stamp desc, index asc) is kept in its record as
// two 64-bit integers (slot 16: next time through hkey_of_double, all ones once finished; slot 15:
// hkey_tie), so the search is integer compares and redux.sync.  Same schedule and records as the
// plain kernel up to the order of the six sums.  Needs at least 3 warps.
template <int T>
__global__ void __launch_bounds__(T, 1)
    hermite_resident_spec_kernel(double *__restrict__ state, int n, double stop_time, double eta,
                                 double eps2, long long max_steps,
                                 long long *__restrict__ nsteps_out,
                                 long long *__restrict__ prof) {
  extern __shared__ double rec[];  // [n][HRES_STRIDE]
  constexpr int NW = T / 32;
  constexpr int NONE = 0x7fffffff;
  static_assert(NW >= 3, "thread 0, a selection warp and at least one summing warp");
  // Roles by warp: 0 = thread 0's corrector, 1 = the search, 2.. = the sums, their sources dealt
  // round robin in chunks of 32.  Warps go to the SM's four schedulers by (warp & 3); with 10 warps
  // every scheduler carries two summing warps, which measured best at N = 1024 (2.04 us per step;
  // keeping thread 0's scheduler free of FP64 work: 2.16, dealing the chunks by scheduler capacity:
  // 2.23, handing them out through a shared counter: 2.5 - the atomic's round trip serialises a
  // warp's chunks).
  constexpr int SELW = 1;
  constexpr int NF = NW - 2;   // summing warps
  __shared__ double red[2][NW][6];
  __shared__ double tgt[2][8];  // summing warps -> thread 0: the next target's predicted state
  // double-buffered by step parity, so that one barrier per step is enough
  __shared__ unsigned long long top[2][2][2];  // warp 1 -> all: (k1, k2) of b1 and of b2
  __shared__ unsigned long long newkey[2][2];  // thread 0 -> all: the stepped body's new (k1, k2)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_force = warp >= 2;
  const int frank = warp - 2;  // rank among the summing warps
  SweepParams P;
  P.eps2 = eps2;
  // prof (NBK_HERMITE_PROF=1, a debugging aid): cycles thread 0 spends [0] in the plain force sum,
  // [1] in the corrector, [2] at the barrier after it; [3] warp 1 searching, [4] warp 2 summing,
  // [5] warp 2 at the barrier; [6] steps taken the plain way
  long long pc[7] = {0, 0, 0, 0, 0, 0, 0};
  long long c0 = 0;
  auto K1 = [&](int j) -> unsigned long long & {
    return *reinterpret_cast<unsigned long long *>(rec + j * HRES_STRIDE + 16);
  };
  auto K2 = [&](int j) -> unsigned long long & {
    return *reinterpret_cast<unsigned long long *>(rec + j * HRES_STRIDE + 15);
  };
  for (int j = tid; j < n; j += T) {
    double *r = rec + j * HRES_STRIDE;
#pragma unroll
    for (int c = 0; c < 16; ++c) r[c] = state[(size_t)j * HB + c];
    const double st = r[15];
    K1(j) = hkey_of_double(r[0] + r[14]);
    K2(j) = hkey_tie(st, j);
  }
  __syncthreads();
  auto less = [](unsigned long long a1, unsigned long long a2, unsigned long long b1,
                 unsigned long long b2) { return a1 < b1 || (a1 == b1 && a2 < b2); };
  // The search belongs to warp 1.  Lane g keeps, in registers, the two smallest keys of "its"
  // bodies j = g, g + 32, g + 64, ... (ca = smallest, cb = second).  A step changes one key (the
  // stepped body's) and hides one body (the one being stepped), so only their two groups are
  // looked at again - cooperatively, one body per lane - and the rest is a reduction over the 32
  // cached pairs: ~1000 cycles instead of a walk over all n keys.
  unsigned long long ca1 = HKEY_NONE, ca2 = HKEY_NONE, cb1 = HKEY_NONE, cb2 = HKEY_NONE;
  auto keep2 = [&](unsigned long long k1, unsigned long long k2, unsigned long long &a1,
                   unsigned long long &a2, unsigned long long &b1, unsigned long long &b2) {
    const bool lt_a = less(k1, k2, a1, a2), lt_b = less(k1, k2, b1, b2);
    b1 = lt_a ? a1 : (lt_b ? k1 : b1);
    b2 = lt_a ? a2 : (lt_b ? k2 : b2);
    a1 = lt_a ? k1 : a1;
    a2 = lt_a ? k2 : a2;
  };
  // warp-wide: the two smallest of the lanes' (a, b) pairs; every lane gets the result
  auto combine2 = [&](unsigned long long a1, unsigned long long a2, unsigned long long b1,
                      unsigned long long b2, unsigned long long (&o)[2][2]) {
    unsigned long long g1 = a1, g2 = a2;
    warp_min_pair(g1, g2);
    // the runner-up: a lane's best if it is not the winner (the tie key holds the index, so
    // keys are unique), else that lane's second
    const bool mine = (a1 == g1 && a2 == g2);
    unsigned long long s1 = mine ? b1 : a1, s2 = mine ? b2 : a2;
    warp_min_pair(s1, s2);
    o[0][0] = g1, o[0][1] = g2, o[1][0] = s1, o[1][1] = s2;
  };
  auto scan_own = [&]() {  // every lane walks its own group (start-up only)
    ca1 = ca2 = cb1 = cb2 = HKEY_NONE;
#pragma unroll 4
    for (int j = lane; j < n; j += 32) keep2(K1(j), K2(j), ca1, ca2, cb1, cb2);
  };
  auto rescan_group = [&](int g, int skip) {  // all lanes look at group g, lane g keeps the result
    unsigned long long a1 = HKEY_NONE, a2 = HKEY_NONE, b1 = HKEY_NONE, b2 = HKEY_NONE;
    for (int j = g + 32 * lane; j < n; j += 32 * 32)
      keep2(j == skip ? HKEY_NONE : K1(j), K2(j), a1, a2, b1, b2);
    unsigned long long o[2][2];
    combine2(a1, a2, b1, b2, o);
    if (lane == g) ca1 = o[0][0], ca2 = o[0][1], cb1 = o[1][0], cb2 = o[1][1];
  };
  auto idx_of = [](unsigned long long k1, unsigned long long k2) {
    return k1 == HKEY_NONE ? NONE : (int)(k2 & ((1u << HKEY_IDX_BITS) - 1));
  };
  // the six sums for target (tg, itarget) over sources != skip, by `nwarps` warps of which this
  // one is number `rank`, into red[buf][rank]
  auto force = [&](int rank, int nwarps, int skip, int itarget, const OpGradVDGradV::Tgt &tg,
                   double nt, int buf) {
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int j = rank * 32 + lane; j < n; j += nwarps * 32) {
      if (j == skip) continue;
      double mj, qj[3], vj[3];
      hermite_predict_fast(rec + j * HRES_STRIDE, nt, mj, qj, vj);
      const double2 src[4] = {make_double2(qj[0], qj[1]), make_double2(qj[2], mj),
                              make_double2(vj[0], vj[1]), make_double2(vj[2], 0.0)};
      OpGradVDGradV::pair<true>(tg, acc, src, src, j == itarget, P);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int c = 0; c < 6; ++c) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], off);
    if (lane == 0)
#pragma unroll
      for (int c = 0; c < 6; ++c) red[buf][rank][c] = acc[c];
  };

  long long steps = 0, stamp = 0;
  bool finishing = false;
  int buf = 0;        // red[buf] holds the sums of the current step
  bool have = false;  // ... speculatively: warps 2.., everything but the interaction with `prev`
  int prev = NONE;    // the body stepped just before
  // the first two bodies of the schedule (warp 1 searches, everybody reads)
  unsigned long long t2[2][2];
  if (warp == SELW) {
    scan_own();
    combine2(ca1, ca2, cb1, cb2, t2);
    if (lane == 0)
      top[1][0][0] = t2[0][0], top[1][0][1] = t2[0][1], top[1][1][0] = t2[1][0], top[1][1][1] = t2[1][1];
  }
  __syncthreads();
  int it = idx_of(top[1][0][0], top[1][0][1]);
  unsigned long long itk1 = top[1][0][0];
  int r = idx_of(top[1][1][0], top[1][1][1]);  // the target after this one, unless `it` repeats
  unsigned long long rk1 = top[1][1][0];
  int par = 0;
  for (;;) {
    if (it == NONE) break;  // every body finished
    const double knt = hkey_to_double(itk1);
    if (!finishing) {
      if (steps >= max_steps) break;
      if (knt > stop_time) finishing = true;  // advance_bodies -> finish_bs (:172-173)
    }
    double *tr = rec + it * HRES_STRIDE;
    const double h = finishing ? __dsub_rn(stop_time, tr[0]) : tr[14];
    const double nt = __dadd_rn(tr[0], h);
    double mt = 0.0, qt[3] = {0, 0, 0}, vt[3] = {0, 0, 0};
    // the target's predicted state: thread 0 needs it for the corrector (the summing warps left
    // it in tgt[] when they prepared this step), everybody when the step is done the plain way
    if (!have) {
      hermite_predict_fast(tr, nt, mt, qt, vt);
    } else if (tid == 0) {
      mt = tgt[buf][0];
#pragma unroll
      for (int c = 0; c < 3; ++c) qt[c] = tgt[buf][1 + c], vt[c] = tgt[buf][4 + c];
    }
    const OpGradVDGradV::Tgt tg = {qt[0], qt[1], qt[2], vt[0], vt[1], vt[2]};
    if (prof) c0 = clock64();
    if (!have) {
      force(warp, NW, NONE, it, tg, nt, buf);
      __syncthreads();
      ++pc[6];
    }
    if (prof) {
      const long long c1 = clock64();
      pc[0] += c1 - c0;
      c0 = c1;
    }
    ++steps;
    ++stamp;
    if (tid == 0) {
      // ---- corrector (hermite.ml:116-123), new time step (hermite.ml:87-93)
      double extra[6] = {0, 0, 0, 0, 0, 0};
      if (have) {  // the one interaction the speculative sums left out: the body stepped last
        double mj, qj[3], vj[3];
        hermite_predict_fast(rec + prev * HRES_STRIDE, nt, mj, qj, vj);
        const double2 src[4] = {make_double2(qj[0], qj[1]), make_double2(qj[2], mj),
                                make_double2(vj[0], vj[1]), make_double2(vj[2], 0.0)};
        OpGradVDGradV::pair<false>(tg, extra, src, src, false, P);
      }
      double fp[3], dfp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double a = extra[c], d = extra[3 + c];
        for (int w = 0; w < (have ? NF : NW); ++w) {
          a += red[buf][w][c];
          d += red[buf][w][3 + c];
        }
        fp[c] = mt * a;  // f = -gsum = m_i * acc
        dfp[c] = mt * d;
      }
      const double m = mt;
      double qn[3], pn[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double q0 = tr[2 + c], p0 = tr[5 + c], f0 = tr[8 + c], df0 = tr[11 + c];
        pn[c] = __dadd_rn(__dadd_rn(p0, __dmul_rn(__dmul_rn(0.5, h), __dadd_rn(f0, fp[c]))),
                          __dmul_rn(__ddiv_rn(__dmul_rn(h, h), 12.0), __dsub_rn(df0, dfp[c])));
        qn[c] = __dadd_rn(
            __dadd_rn(q0, __dmul_rn(__ddiv_rn(__dmul_rn(0.5, h), m), __dadd_rn(p0, pn[c]))),
            __dmul_rn(__ddiv_rn(__dmul_rn(h, h), __dmul_rn(12.0, m)), __dsub_rn(f0, fp[c])));
      }
      const double dx = qn[0] - qt[0], dy = qn[1] - qt[1], dz = qn[2] - qt[2];
      const double rr = __dsqrt_rn(
          __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
      const double fn = __dsqrt_rn(__dadd_rn(
          __dadd_rn(__dmul_rn(fp[0], fp[0]), __dmul_rn(fp[1], fp[1])), __dmul_rn(fp[2], fp[2])));
      double hn;
      if (rr == 0.0) {
        hn = __dmul_rn(h, 1.189207115002721);  // 2 ** 0.25
      } else {
        const double x =
            __ddiv_rn(__dmul_rn(__dmul_rn(eta, __dmul_rn(h, h)), fn), __dmul_rn(m, rr));
        hn = __dmul_rn(h, __dsqrt_rn(__dsqrt_rn(x)));  // x ** 0.25
      }
      tr[0] = nt;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        tr[2 + c] = qn[c];
        tr[5 + c] = pn[c];
        tr[8 + c] = fp[c];
        tr[11 + c] = dfp[c];
      }
      tr[14] = hn;
      const unsigned long long nk1 = finishing ? HKEY_NONE : hkey_of_double(nt + hn);
      const unsigned long long nk2 = hkey_tie((double)stamp, it);
      K1(it) = nk1;
      K2(it) = nk2;
      newkey[par][0] = nk1;
      newkey[par][1] = nk2;
      if (prof) {
        const long long c1 = clock64();
        pc[1] += c1 - c0;
        c0 = c1;
      }
    } else if (warp == SELW) {
      // ---- meanwhile: the two earliest bodies other than `it` (whose record is in flux)
      if (prev != NONE && (prev & 31) != (it & 31)) rescan_group(prev & 31, it);
      rescan_group(it & 31, it);
      combine2(ca1, ca2, cb1, cb2, t2);
      if (lane == 0)
        top[par][0][0] = t2[0][0], top[par][0][1] = t2[0][1], top[par][1][0] = t2[1][0],
        top[par][1][1] = t2[1][1];
      if (prof) {
        const long long c1 = clock64();
        pc[3] += c1 - c0;
        c0 = c1;
      }
    } else if (is_force && r != NONE) {
      // ---- meanwhile: the sums of the next step's target r, without `it`
      const bool fin_r = finishing || hkey_to_double(rk1) > stop_time;
      const double *rr_ = rec + r * HRES_STRIDE;
      const double hr = fin_r ? __dsub_rn(stop_time, rr_[0]) : rr_[14];
      const double ntr = __dadd_rn(rr_[0], hr);
      double mr, qr[3], vr[3];
      hermite_predict_fast(rr_, ntr, mr, qr, vr);
      const OpGradVDGradV::Tgt tgr = {qr[0], qr[1], qr[2], vr[0], vr[1], vr[2]};
      if (frank == 0 && lane == 0) {
        tgt[buf ^ 1][0] = mr;
#pragma unroll
        for (int c = 0; c < 3; ++c) tgt[buf ^ 1][1 + c] = qr[c], tgt[buf ^ 1][4 + c] = vr[c];
      }
      force(frank, NF, it, r, tgr, ntr, buf ^ 1);
      if (prof) {
        const long long c1 = clock64();
        pc[4] += c1 - c0;
        c0 = c1;
      }
    }
    __syncthreads();
    if (prof) {
      const long long c1 = clock64();
      pc[tid == 0 ? 2 : 5] += c1 - c0;
    }
    // ---- who is next, and who after that.  b1 (== r) and b2 are the earliest bodies other than
    // `it`; `it` re-enters with its new key unless it is finished.
    const unsigned long long b1k1 = top[par][0][0], b1k2 = top[par][0][1], b2k1 = top[par][1][0],
                             b2k2 = top[par][1][1];
    const unsigned long long nk1 = newkey[par][0], nk2 = newkey[par][1];
    par ^= 1;
    const bool again = nk1 != HKEY_NONE && (b1k1 == HKEY_NONE || less(nk1, nk2, b1k1, b1k2));
    prev = it;
    if (again) {
      // `it` stays; the body after it is b1
      itk1 = nk1;
      have = false;
      r = idx_of(b1k1, b1k2);
      rk1 = b1k1;
    } else {
      it = idx_of(b1k1, b1k2);
      itk1 = b1k1;
      have = (it != NONE);
      buf ^= 1;
      // excluding b1, the earliest is b2 or the re-inserted body
      if (nk1 != HKEY_NONE && less(nk1, nk2, b2k1, b2k2)) {
        r = prev;
        rk1 = nk1;
      } else {
        r = idx_of(b2k1, b2k2);
        rk1 = b2k1;
      }
    }
  }
  __syncthreads();
  for (int j = tid; j < n; j += T) {
    double *rj = rec + j * HRES_STRIDE;
    // slot 15 back to the re-insertion stamp the plain kernel keeps there
    rj[15] = (double)(0x7ffffffffffull - (K2(j) >> HKEY_IDX_BITS));
#pragma unroll
    for (int c = 0; c < 16; ++c) state[(size_t)j * HB + c] = rj[c];
  }
  if (tid == 0) *nsteps_out = steps;
  if (prof) {
    if (tid == 0) prof[0] = pc[0], prof[1] = pc[1], prof[2] = pc[2], prof[6] = pc[6];
    if (tid == 32 * SELW) prof[3] = pc[3];
    if (is_force && frank == 0 && lane == 0) prof[4] = pc[4], prof[5] = pc[5];
  }
}

// ---- the same stepping loop on a thread-block CLUSTER (1536 < N <= 12288) ----------------------
// Eight CTAs (one cluster, eight SMs) each keep an eighth of the Hermite.b array in their own
// shared memory and talk through distributed shared memory (DSMEM) - two cluster barriers per
// step instead of a kernel launch:
//   post:  every CTA finds its earliest unfinished body and stores that body's key and whole
//          record into slot [rank] of the candidate table of ALL eight CTAs          (DSMEM)
//   sync A
//   pick:  every CTA picks the same winner from its copy of the table (so it already holds the
//          stepped body's record), predicts its own bodies, sums its part of the 1 x (N-1)
//          acc+jerk and stores the six partial sums into the OWNER CTA's table      (DSMEM)
//   sync B
//   owner: adds the eight partials in rank order, corrector + new time step, updates its record.
constexpr int HCL_CTAS = 8;
constexpr int HCL_THREADS = 256;
constexpr int HCL_PER_CTA = 1536;
constexpr int HCL_MAX_N = HCL_CTAS * HCL_PER_CTA;
constexpr int HCL_CAND = 20;  // nt, stamp, index, then the 16-double record, pad

__global__ void __cluster_dims__(HCL_CTAS, 1, 1) __launch_bounds__(HCL_THREADS, 1)
    hermite_cluster_kernel(double *__restrict__ state, int n, double stop_time, double eta,
                           double eps2, long long max_steps, long long *__restrict__ nsteps_out) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  // [nloc][HRES_STRIDE]; slots 15 / 16 hold the body's rank as two integer keys (hkey_tie of
  // stamp and global index; hkey_of_double of the next time, all ones once finished)
  extern __shared__ double rec[];
  __shared__ double cand[HCL_CTAS][HCL_CAND];  // key1, key2 (bit patterns), -, then the record
  __shared__ double part[HCL_CTAS][6];
  __shared__ double red[HCL_THREADS / 32][6];
  __shared__ unsigned long long loc_key[2];  // warp 0 -> CTA: this CTA's earliest body
  constexpr int NW = HCL_THREADS / 32;
  constexpr int NONE = 0x7fffffff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + HCL_CTAS - 1) / HCL_CTAS;
  const int j_lo = min(rank * per, n), j_hi = min(j_lo + per, n), nloc = j_hi - j_lo;
  SweepParams P;
  P.eps2 = eps2;
  for (int l = tid; l < nloc; l += HCL_THREADS) {
#pragma unroll
    for (int c = 0; c < 16; ++c) rec[l * HRES_STRIDE + c] = state[(size_t)(j_lo + l) * HB + c];
  }
  auto K1 = [&](int l) -> unsigned long long & {
    return *reinterpret_cast<unsigned long long *>(rec + l * HRES_STRIDE + 16);
  };
  auto K2 = [&](int l) -> unsigned long long & {
    return *reinterpret_cast<unsigned long long *>(rec + l * HRES_STRIDE + 15);
  };
  for (int l = tid; l < nloc; l += HCL_THREADS) {
    const double *r = rec + l * HRES_STRIDE;
    const double st = r[15];
    K1(l) = hkey_of_double(r[0] + r[14]);
    K2(l) = hkey_tie(st, j_lo + l);
  }
  __syncthreads();
  long long steps = 0, stamp = 0;
  bool finishing = false;
  auto less = [](unsigned long long a1, unsigned long long a2, unsigned long long b1,
                 unsigned long long b2) { return a1 < b1 || (a1 == b1 && a2 < b2); };
  // The CTA's earliest body (as in hermite_resident_spec_kernel): lane g of warp 0 keeps the
  // smallest key of the local bodies l = g, g + 32, ...; a step changes one key in one CTA, whose
  // warp 0 re-examines that one class, and the candidate is a redux.sync reduction of the 32
  // cached keys - instead of every thread walking the records and two rounds of shuffles.
  unsigned long long c1 = HKEY_NONE, c2 = HKEY_NONE;
  if (warp == 0)
    for (int l = lane; l < nloc; l += 32) {
      const unsigned long long k1 = K1(l), k2 = K2(l);
      if (less(k1, k2, c1, c2)) c1 = k1, c2 = k2;
    }
  int changed = -1;  // local index of the body whose key changed in the last step (owner only)
  for (;;) {
    // ---- post: this CTA's earliest body (next time asc, stamp desc, global index asc)
    if (warp == 0) {
      if (changed >= 0) {  // all lanes look at the class of the body that was stepped
        const int g = changed & 31;
        unsigned long long a1 = HKEY_NONE, a2 = HKEY_NONE;
        for (int l = g + 32 * lane; l < nloc; l += 32 * 32) {
          const unsigned long long k1 = K1(l), k2 = K2(l);
          if (less(k1, k2, a1, a2)) a1 = k1, a2 = k2;
        }
        warp_min_pair(a1, a2);
        if (lane == g) c1 = a1, c2 = a2;
      }
      unsigned long long b1 = c1, b2 = c2;
      warp_min_pair(b1, b2);
      if (lane == 0) loc_key[0] = b1, loc_key[1] = b2;
    }
    __syncthreads();
    {
      const unsigned long long k1 = loc_key[0], k2 = loc_key[1];
      const int kidx = k1 == HKEY_NONE ? NONE : (int)(k2 & ((1u << HKEY_IDX_BITS) - 1));
      if (tid < HCL_CTAS * HCL_CAND) {
        const int dst = tid / HCL_CAND, slot = tid % HCL_CAND;
        double v = 0.0;
        if (slot == 0) v = __longlong_as_double((long long)k1);
        else if (slot == 1) v = __longlong_as_double((long long)k2);
        else if (slot >= 3 && slot < 19 && kidx != NONE)
          v = rec[(kidx - j_lo) * HRES_STRIDE + (slot - 3)];
        double *remote = cluster.map_shared_rank(&cand[0][0], dst);
        remote[rank * HCL_CAND + slot] = v;
      }
    }
    cluster.sync();  // A
    // ---- pick: the same winner everywhere
    int it = NONE, owner = 0;
    {
      unsigned long long b1 = HKEY_NONE, b2 = HKEY_NONE;
      for (int r = 0; r < HCL_CTAS; ++r) {
        const unsigned long long k1 = (unsigned long long)__double_as_longlong(cand[r][0]),
                                 k2 = (unsigned long long)__double_as_longlong(cand[r][1]);
        if (less(k1, k2, b1, b2)) {
          b1 = k1;
          b2 = k2;
          owner = r;
        }
      }
      if (b1 == HKEY_NONE) break;  // every body finished (uniform over the cluster)
      it = (int)(b2 & ((1u << HKEY_IDX_BITS) - 1));
      if (!finishing) {
        if (steps >= max_steps) break;
        if (hkey_to_double(b1) > stop_time) finishing = true;
      }
    }
    const double *tr = &cand[owner][3];
    const double h = finishing ? __dsub_rn(stop_time, tr[0]) : tr[14];
    const double nt = __dadd_rn(tr[0], h);
    double mt, qt[3], vt[3];
    hermite_predict_fast(tr, nt, mt, qt, vt);
    const OpGradVDGradV::Tgt tg = {qt[0], qt[1], qt[2], vt[0], vt[1], vt[2]};
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int l = tid; l < nloc; l += HCL_THREADS) {
      double mj, qj[3], vj[3];
      hermite_predict_fast(rec + l * HRES_STRIDE, nt, mj, qj, vj);
      const double2 src[4] = {make_double2(qj[0], qj[1]), make_double2(qj[2], mj),
                              make_double2(vj[0], vj[1]), make_double2(vj[2], 0.0)};
      OpGradVDGradV::pair<true>(tg, acc, src, src, (j_lo + l) == it, P);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int c = 0; c < 6; ++c) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], off);
    if (lane == 0)
#pragma unroll
      for (int c = 0; c < 6; ++c) red[warp][c] = acc[c];
    __syncthreads();
    if (tid < 6) {
      double a = red[0][tid];
      for (int w = 1; w < NW; ++w) a += red[w][tid];
      double *remote = cluster.map_shared_rank(&part[0][0], owner);
      remote[rank * 6 + tid] = a;
    }
    cluster.sync();  // B
    ++steps;
    ++stamp;
    // ---- owner: corrector (hermite.ml:116-123) and new time step (hermite.ml:87-93)
    if (rank == owner && tid == 0) {
      double fp[3], dfp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double a = 0.0, d = 0.0;
        for (int r = 0; r < HCL_CTAS; ++r) {
          a += part[r][c];
          d += part[r][3 + c];
        }
        fp[c] = mt * a;
        dfp[c] = mt * d;
      }
      double *wr = rec + (it - j_lo) * HRES_STRIDE;
      const double m = mt;
      double qn[3], pn[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double q0 = tr[2 + c], p0 = tr[5 + c], f0 = tr[8 + c], df0 = tr[11 + c];
        pn[c] = __dadd_rn(__dadd_rn(p0, __dmul_rn(__dmul_rn(0.5, h), __dadd_rn(f0, fp[c]))),
                          __dmul_rn(__ddiv_rn(__dmul_rn(h, h), 12.0), __dsub_rn(df0, dfp[c])));
        qn[c] = __dadd_rn(
            __dadd_rn(q0, __dmul_rn(__ddiv_rn(__dmul_rn(0.5, h), m), __dadd_rn(p0, pn[c]))),
            __dmul_rn(__ddiv_rn(__dmul_rn(h, h), __dmul_rn(12.0, m)), __dsub_rn(f0, fp[c])));
      }
      const double dx = qn[0] - qt[0], dy = qn[1] - qt[1], dz = qn[2] - qt[2];
      const double rr = __dsqrt_rn(
          __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
      const double fn = __dsqrt_rn(__dadd_rn(
          __dadd_rn(__dmul_rn(fp[0], fp[0]), __dmul_rn(fp[1], fp[1])), __dmul_rn(fp[2], fp[2])));
      double hn;
      if (rr == 0.0) {
        hn = __dmul_rn(h, 1.189207115002721);
      } else {
        const double x =
            __ddiv_rn(__dmul_rn(__dmul_rn(eta, __dmul_rn(h, h)), fn), __dmul_rn(m, rr));
        hn = __dmul_rn(h, __dsqrt_rn(__dsqrt_rn(x)));
      }
      wr[0] = nt;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        wr[2 + c] = qn[c];
        wr[5 + c] = pn[c];
        wr[8 + c] = fp[c];
        wr[11 + c] = dfp[c];
      }
      wr[14] = hn;
      K1(it - j_lo) = finishing ? HKEY_NONE : hkey_of_double(nt + hn);
      K2(it - j_lo) = hkey_tie((double)stamp, it);
    }
    changed = rank == owner ? it - j_lo : -1;
    __syncthreads();  // the owner's next search must see the update; uniform in every CTA
  }
  cluster.sync();  // nobody leaves while a peer may still write into its shared memory
  for (int l = tid; l < nloc; l += HCL_THREADS) {
    double *rl = rec + l * HRES_STRIDE;
    rl[15] = (double)(0x7ffffffffffull - (K2(l) >> HKEY_IDX_BITS));  // back to the stamp
#pragma unroll
    for (int c = 0; c < 16; ++c) state[(size_t)(j_lo + l) * HB + c] = rl[c];
  }
  if (rank == 0 && tid == 0) *nsteps_out = steps;
}

// ---- the cluster loop, software-pipelined ------------------------------------------------------
// hermite_cluster_kernel serialises search -> barrier -> sum -> barrier -> corrector, and the
// corrector (one thread of one CTA, ~1900 cycles) keeps eight SMs waiting.  Here, as in
// hermite_resident_spec_kernel, a step k with target i_k already knows the NEXT target r_k (the
// earliest body other than i_k), so three things run side by side and ONE cluster barrier ends
// the step:
//   thread 0 of i_k's owner   sums of step k (the eight partials + the one interaction with the
//                             body stepped before), corrector, new keys; the new record goes to
//                             all CTAs (DSMEM)
//   warp 1 of every CTA       the CTA's two earliest bodies other than i_k (cached class top-2,
//                             integer keys) with their records -> all CTAs (DSMEM)
//   warps 2.. of every CTA    predict the local bodies to r_k's time, local part of the
//                             1 x (N-2) interactions without i_k, six partials -> r_k's owner
//   cluster barrier
//   everybody                 the two smallest keys among the sixteen candidates and the re-inserted
//                             i_k: the smallest is the next target (r_k, or i_k again - ties go
//                             to the re-inserted body; then that step is done the plain way, all
//                             warps, one more barrier), the second smallest the target after it.
constexpr int HCLS_THREADS = 320;

__global__ void __cluster_dims__(HCL_CTAS, 1, 1) __launch_bounds__(HCLS_THREADS, 1)
    hermite_cluster_spec_kernel(double *__restrict__ state, int n, double stop_time, double eta,
                                double eps2, long long max_steps,
                                long long *__restrict__ nsteps_out,
                                long long *__restrict__ prof) {
  // prof (NBK_HERMITE_PROF=1, rank 0 only): cycles [0] loop top + plain path, [1] thread 0 corrector
  // (when rank 0 owns the target), [2] warp 1 search + post, [3] warp 2 local sums + post,
  // [4] barrier as seen by thread 0, [5] pick + decision, [6] steps owned by rank 0
  long long pc[7] = {0, 0, 0, 0, 0, 0, 0};
  long long c0 = prof ? clock64() : 0, c1;
#define HCLS_TICK(slot)            \
  if (prof) {                      \
    c1 = clock64();                \
    pc[slot] += c1 - c0;           \
    c0 = c1;                       \
  }
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  constexpr int T = HCLS_THREADS, NW = T / 32, NF = NW - 2;
  constexpr int NONE = 0x7fffffff;
  constexpr int NCAND = 2 * HCL_CTAS + 1;  // two per CTA and the re-inserted body
  extern __shared__ double rec[];  // [nloc][HRES_STRIDE]; slots 15 / 16 = the integer keys
  // every table is double-buffered by step parity: a step rewrites what was read two steps ago.
  // A candidate row: key1, key2 (bit patterns), -, record[16], pad.
  __shared__ double cand[2][NCAND][HCL_CAND];  // row 2 c + s: CTA c's s-th; row 16: the stepped body
  __shared__ double part[2][HCL_CTAS][6];      // partial sums, at the target's owner
  __shared__ double tgtp[2][8];                // the next target's predicted state (local copy)
  __shared__ double red[NW][6];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + HCL_CTAS - 1) / HCL_CTAS;
  const int j_lo = min(rank * per, n), j_hi = min(j_lo + per, n), nloc = j_hi - j_lo;
  SweepParams P;
  P.eps2 = eps2;
  auto K1 = [&](int l) -> unsigned long long & {
    return *reinterpret_cast<unsigned long long *>(rec + l * HRES_STRIDE + 16);
  };
  auto K2 = [&](int l) -> unsigned long long & {
    return *reinterpret_cast<unsigned long long *>(rec + l * HRES_STRIDE + 15);
  };
  for (int l = tid; l < nloc; l += T) {
    double *r = rec + l * HRES_STRIDE;
#pragma unroll
    for (int c = 0; c < 16; ++c) r[c] = state[(size_t)(j_lo + l) * HB + c];
    const double st = r[15];
    K1(l) = hkey_of_double(r[0] + r[14]);
    K2(l) = hkey_tie(st, j_lo + l);
  }
  __syncthreads();
  auto less = [](unsigned long long a1, unsigned long long a2, unsigned long long b1,
                 unsigned long long b2) { return a1 < b1 || (a1 == b1 && a2 < b2); };
  auto key_of = [](const double *c, int w) {
    return (unsigned long long)__double_as_longlong(c[w]);
  };
  auto idx_of = [](unsigned long long k1, unsigned long long k2) {
    return k1 == HKEY_NONE ? NONE : (int)(k2 & ((1u << HKEY_IDX_BITS) - 1));
  };
  auto keep2 = [&](unsigned long long k1, unsigned long long k2, unsigned long long &a1,
                   unsigned long long &a2, unsigned long long &b1, unsigned long long &b2) {
    const bool lt_a = less(k1, k2, a1, a2), lt_b = less(k1, k2, b1, b2);
    b1 = lt_a ? a1 : (lt_b ? k1 : b1);
    b2 = lt_a ? a2 : (lt_b ? k2 : b2);
    a1 = lt_a ? k1 : a1;
    a2 = lt_a ? k2 : a2;
  };
  // warp-wide: the two smallest of the lanes' (a, b) pairs; every lane gets the result
  auto combine2 = [&](unsigned long long a1, unsigned long long a2, unsigned long long b1,
                      unsigned long long b2, unsigned long long (&o)[2][2]) {
    unsigned long long g1 = a1, g2 = a2;
    warp_min_pair(g1, g2);
    const bool mine = (a1 == g1 && a2 == g2);  // keys are unique (the tie key holds the index)
    unsigned long long s1 = mine ? b1 : a1, s2 = mine ? b2 : a2;
    warp_min_pair(s1, s2);
    o[0][0] = g1, o[0][1] = g2, o[1][0] = s1, o[1][1] = s2;
  };
  // warp 1: this CTA's two earliest bodies other than the local body `hide` -> rows [2 rank],
  // [2 rank + 1] of every CTA's cand[pb].  Up to 384 local bodies: a plain walk over the keys
  // (nloc / 32 per lane, branch-free, ~100 cycles each) and one warp-wide top-2.  Beyond: lane g
  // caches the two smallest keys of the local bodies g, g + 32, ... and a step re-examines only
  // the class of the body stepped before (`changed`: its key changed, and it was hidden then) and
  // the class of `hide` - three warp-wide top-2, ~560 cycles each (redux.sync costs ~70 cycles).
  const bool cached = nloc > 384;
  unsigned long long ca1 = HKEY_NONE, ca2 = HKEY_NONE, cb1 = HKEY_NONE, cb2 = HKEY_NONE;
  auto rescan = [&](int g, int hide) {  // all lanes look at class g, lane g keeps the result
    unsigned long long a1 = HKEY_NONE, a2 = HKEY_NONE, b1 = HKEY_NONE, b2 = HKEY_NONE;
    for (int l = g + 32 * lane; l < nloc; l += 32 * 32)
      keep2(l == hide ? HKEY_NONE : K1(l), K2(l), a1, a2, b1, b2);
    unsigned long long o[2][2];
    combine2(a1, a2, b1, b2, o);
    if (lane == g) ca1 = o[0][0], ca2 = o[0][1], cb1 = o[1][0], cb2 = o[1][1];
  };
  auto post_candidates = [&](int pb, int hide, int changed) {
    unsigned long long a1 = HKEY_NONE, a2 = HKEY_NONE, b1 = HKEY_NONE, b2 = HKEY_NONE;
    if (cached) {
      if (changed >= 0) rescan(changed & 31, hide);
      if (hide >= 0 && (changed < 0 || (changed & 31) != (hide & 31))) rescan(hide & 31, hide);
      a1 = ca1, a2 = ca2, b1 = cb1, b2 = cb2;
    } else {
#pragma unroll 4
      for (int l = lane; l < nloc; l += 32) keep2(l == hide ? HKEY_NONE : K1(l), K2(l), a1, a2, b1, b2);
    }
    unsigned long long o[2][2];
    combine2(a1, a2, b1, b2, o);
    // 2 x 20 values, computed once, stored into all eight tables
    for (int v = lane; v < 2 * HCL_CAND; v += 32) {
      const int s = v / HCL_CAND, slot = v - s * HCL_CAND;
      const int kl = o[s][0] == HKEY_NONE ? -1 : (int)(o[s][1] & ((1u << HKEY_IDX_BITS) - 1)) - j_lo;
      double x = 0.0;
      if (slot == 0) x = __longlong_as_double((long long)o[s][0]);
      else if (slot == 1) x = __longlong_as_double((long long)o[s][1]);
      else if (slot >= 3 && slot < 19 && kl >= 0) x = rec[kl * HRES_STRIDE + (slot - 3)];
      const size_t at = ((size_t)pb * NCAND + 2 * rank + s) * HCL_CAND + slot;
#pragma unroll
      for (int dst = 0; dst < HCL_CTAS; ++dst) cluster.map_shared_rank(&cand[0][0][0], dst)[at] = x;
    }
  };
  // every warp: the two smallest keys of cand[pb]'s rows (lane x looks at row x) -> row numbers
  auto pick2 = [&](int pb, int nrows, int &row1, int &row2) {
    unsigned long long a1 = HKEY_NONE, a2 = HKEY_NONE;
    if (lane < nrows) a1 = key_of(cand[pb][lane], 0), a2 = key_of(cand[pb][lane], 1);
    unsigned long long g1 = a1, g2 = a2;
    warp_min_pair(g1, g2);
    const bool mine = lane < nrows && a1 == g1 && a2 == g2 && g1 != HKEY_NONE;
    unsigned long long s1 = mine ? HKEY_NONE : a1, s2 = mine ? HKEY_NONE : a2;
    warp_min_pair(s1, s2);
    const bool second = lane < nrows && !mine && a1 == s1 && a2 == s2 && s1 != HKEY_NONE;
    const unsigned b1 = __ballot_sync(0xffffffffu, mine), b2 = __ballot_sync(0xffffffffu, second);
    row1 = b1 ? __ffs(b1) - 1 : -1;
    row2 = b2 ? __ffs(b2) - 1 : -1;
  };
  // warps [w0, NW): the local part of the six sums for target tg (global index itarget) at time nt,
  // leaving out the local body `hide`; then one thread per component posts the CTA's partial into
  // row [rank] of part[pb] at CTA `own`.  `sync` = a barrier of exactly those warps.
  auto local_force = [&](int w0, int hide, int itarget, const OpGradVDGradV::Tgt &tg, double nt,
                         int pb, int own, auto sync) {
    const int nthr = (NW - w0) * 32, t0 = tid - w0 * 32;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int l = t0; l < nloc; l += nthr) {
      if (l == hide) continue;
      double mj, qj[3], vj[3];
      hermite_predict_fast(rec + l * HRES_STRIDE, nt, mj, qj, vj);
      const double2 src[4] = {make_double2(qj[0], qj[1]), make_double2(qj[2], mj),
                              make_double2(vj[0], vj[1]), make_double2(vj[2], 0.0)};
      OpGradVDGradV::pair<true>(tg, acc, src, src, (j_lo + l) == itarget, P);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int c = 0; c < 6; ++c) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], off);
    if (lane == 0)
#pragma unroll
      for (int c = 0; c < 6; ++c) red[warp][c] = acc[c];
    sync();
    if (t0 < 6) {
      double a = red[w0][t0];
      for (int w = w0 + 1; w < NW; ++w) a += red[w][t0];
      double *remote = cluster.map_shared_rank(&part[0][0][0], own);
      remote[(pb * HCL_CTAS + rank) * 6 + t0] = a;
    }
  };
  auto sync_all = [] { __syncthreads(); };
  auto sync_force = [] { asm volatile("bar.sync 2, %0;" ::"r"(NF * 32) : "memory"); };

  long long steps = 0, stamp = 0;
  bool finishing = false;
  // ---- the first two bodies of the schedule
  if (warp == 1) {
    if (cached) {
#pragma unroll 4
      for (int l = lane; l < nloc; l += 32) keep2(K1(l), K2(l), ca1, ca2, cb1, cb2);
    }
    post_candidates(1, -1, -1);
  }
  cluster.sync();
  int row1, row2;
  pick2(1, 2 * HCL_CTAS, row1, row2);
  const double *cur = row1 >= 0 ? &cand[1][row1][0] : nullptr;  // the current target: keys, -, record
  const double *nxt = row2 >= 0 ? &cand[1][row2][0] : nullptr;  // the target after it
  int par = 0;      // buffer of this step's candidate table
  int pb = 0;       // buffer of this step's partial sums / predicted target
  bool have = false;
  const double *prevrec = nullptr;  // the body stepped before (its new record), for the extra pair
  int changed = -1;  // local index of the body stepped before (its key changed), -1 if not local
  for (;;) {
    if (cur == nullptr) break;  // every body finished (uniform over the cluster)
    const unsigned long long itk1 = key_of(cur, 0);
    const int it = idx_of(itk1, key_of(cur, 1)), owner = it / per;
    if (!finishing) {
      if (steps >= max_steps) break;
      if (hkey_to_double(itk1) > stop_time) finishing = true;
    }
    const double *tr = cur + 3;
    const double h = finishing ? __dsub_rn(stop_time, tr[0]) : tr[14];
    const double nt = __dadd_rn(tr[0], h);
    const int it_loc = rank == owner ? it - j_lo : -1;
    double mt = 0.0, qt[3] = {0, 0, 0}, vt[3] = {0, 0, 0};
    if (!have) {
      // the plain way: everybody sums its local part for `it`, one more cluster barrier
      hermite_predict_fast(tr, nt, mt, qt, vt);
      const OpGradVDGradV::Tgt tg = {qt[0], qt[1], qt[2], vt[0], vt[1], vt[2]};
      local_force(0, -1, it, tg, nt, pb, owner, sync_all);
      cluster.sync();
    } else if (tid == 0 && rank == owner) {
      mt = tgtp[pb][0];
#pragma unroll
      for (int c = 0; c < 3; ++c) qt[c] = tgtp[pb][1 + c], vt[c] = tgtp[pb][4 + c];
    }
    ++steps;
    ++stamp;
    const int r = nxt ? idx_of(key_of(nxt, 0), key_of(nxt, 1)) : NONE;
    HCLS_TICK(0)
    if (warp == 0) {
      if (tid == 0 && rank == owner) {
        // ---- corrector (hermite.ml:116-123), new time step (hermite.ml:87-93)
        const OpGradVDGradV::Tgt tg = {qt[0], qt[1], qt[2], vt[0], vt[1], vt[2]};
        double extra[6] = {0, 0, 0, 0, 0, 0};
        if (have) {  // the interaction the speculative sums left out: the body stepped last
          double mj, qj[3], vj[3];
          hermite_predict_fast(prevrec + 3, nt, mj, qj, vj);
          const double2 src[4] = {make_double2(qj[0], qj[1]), make_double2(qj[2], mj),
                                  make_double2(vj[0], vj[1]), make_double2(vj[2], 0.0)};
          OpGradVDGradV::pair<false>(tg, extra, src, src, false, P);
        }
        double fp[3], dfp[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double a = extra[c], d = extra[3 + c];
          for (int q = 0; q < HCL_CTAS; ++q) {
            a += part[pb][q][c];
            d += part[pb][q][3 + c];
          }
          fp[c] = mt * a;
          dfp[c] = mt * d;
        }
        double *wr = rec + it_loc * HRES_STRIDE;
        const double m = mt;
        double qn[3], pn[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double q0 = tr[2 + c], p0 = tr[5 + c], f0 = tr[8 + c], df0 = tr[11 + c];
          pn[c] = __dadd_rn(__dadd_rn(p0, __dmul_rn(__dmul_rn(0.5, h), __dadd_rn(f0, fp[c]))),
                            __dmul_rn(__ddiv_rn(__dmul_rn(h, h), 12.0), __dsub_rn(df0, dfp[c])));
          qn[c] = __dadd_rn(
              __dadd_rn(q0, __dmul_rn(__ddiv_rn(__dmul_rn(0.5, h), m), __dadd_rn(p0, pn[c]))),
              __dmul_rn(__ddiv_rn(__dmul_rn(h, h), __dmul_rn(12.0, m)), __dsub_rn(f0, fp[c])));
        }
        const double dx = qn[0] - qt[0], dy = qn[1] - qt[1], dz = qn[2] - qt[2];
        const double rr = __dsqrt_rn(
            __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        const double fn = __dsqrt_rn(__dadd_rn(
            __dadd_rn(__dmul_rn(fp[0], fp[0]), __dmul_rn(fp[1], fp[1])), __dmul_rn(fp[2], fp[2])));
        double hn;
        if (rr == 0.0) {
          hn = __dmul_rn(h, 1.189207115002721);
        } else {
          const double x =
              __ddiv_rn(__dmul_rn(__dmul_rn(eta, __dmul_rn(h, h)), fn), __dmul_rn(m, rr));
          hn = __dmul_rn(h, __dsqrt_rn(__dsqrt_rn(x)));
        }
        wr[0] = nt;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          wr[2 + c] = qn[c];
          wr[5 + c] = pn[c];
          wr[8 + c] = fp[c];
          wr[11 + c] = dfp[c];
        }
        wr[14] = hn;
        K1(it_loc) = finishing ? HKEY_NONE : hkey_of_double(nt + hn);
        K2(it_loc) = hkey_tie((double)stamp, it);
      }
      __syncwarp();
      if (rank == owner && lane < 19) {
        // the new keys and record -> row 16 of every CTA's candidate table
        const double *wr = rec + it_loc * HRES_STRIDE;
        double v = 0.0;
        if (lane == 0) v = wr[16];
        else if (lane == 1) v = wr[15];
        else if (lane >= 3) v = wr[lane - 3];
        for (int dst = 0; dst < HCL_CTAS; ++dst) {
          double *remote = cluster.map_shared_rank(&cand[0][0][0], dst);
          remote[((size_t)par * NCAND + 2 * HCL_CTAS) * HCL_CAND + lane] = v;
        }
      }
      if (rank == owner) {
        HCLS_TICK(1)
        ++pc[6];
      }
    } else if (warp == 1) {
      // ---- this CTA's two earliest bodies other than `it`
      post_candidates(par, it_loc, changed);
      HCLS_TICK(2)
    } else if (r != NONE) {
      // ---- the local part of the next target's sums, without `it`
      const double *rr_ = nxt + 3;
      const bool fin_r = finishing || hkey_to_double(key_of(nxt, 0)) > stop_time;
      const double hr = fin_r ? __dsub_rn(stop_time, rr_[0]) : rr_[14];
      const double ntr = __dadd_rn(rr_[0], hr);
      double mr, qr[3], vr[3];
      hermite_predict_fast(rr_, ntr, mr, qr, vr);
      const OpGradVDGradV::Tgt tgr = {qr[0], qr[1], qr[2], vr[0], vr[1], vr[2]};
      if (tid == 64) {
        tgtp[pb ^ 1][0] = mr;
#pragma unroll
        for (int c = 0; c < 3; ++c) tgtp[pb ^ 1][1 + c] = qr[c], tgtp[pb ^ 1][4 + c] = vr[c];
      }
      local_force(2, it_loc, r, tgr, ntr, pb ^ 1, r / per, sync_force);
      HCLS_TICK(3)
    }
    cluster.sync();
    HCLS_TICK(4)
    // ---- who is next, and who after that (uniform over the cluster): the two smallest keys among
    // the sixteen candidates and the re-inserted body (row 16; all ones when it is finished)
    pick2(par, NCAND, row1, row2);
    changed = it_loc;
    const bool again = row1 == 2 * HCL_CTAS;
    prevrec = &cand[par][2 * HCL_CTAS][0];
    cur = row1 >= 0 ? &cand[par][row1][0] : nullptr;
    nxt = row2 >= 0 ? &cand[par][row2][0] : nullptr;
    // the speculative sums are for r: usable iff r is indeed next
    have = !again && cur != nullptr && idx_of(key_of(cur, 0), key_of(cur, 1)) == r;
    if (have) pb ^= 1;
    par ^= 1;
    HCLS_TICK(5)
  }
  cluster.sync();  // nobody leaves while a peer may still write into its shared memory
  for (int l = tid; l < nloc; l += T) {
    double *rl = rec + l * HRES_STRIDE;
    rl[15] = (double)(0x7ffffffffffull - (K2(l) >> HKEY_IDX_BITS));  // back to the stamp
#pragma unroll
    for (int c = 0; c < 16; ++c) state[(size_t)(j_lo + l) * HB + c] = rl[c];
  }
  if (rank == 0 && tid == 0) *nsteps_out = steps;
  if (prof && rank == 0) {
    if (tid == 0) prof[0] = pc[0], prof[1] = pc[1], prof[4] = pc[4], prof[5] = pc[5], prof[6] = pc[6];
    if (tid == 32) prof[2] = pc[2];
    if (tid == 64) prof[3] = pc[3];
  }
#undef HCLS_TICK
}

// ---- the same stepping loop on the WHOLE GRID (12288 < N <= HGR_MAX_N) --------------------------
// One CTA per SM (cooperative launch), each keeping n / G bodies in its shared memory; what the
// cluster variant passes through distributed shared memory goes through L2 here, with two grid
// barriers per step (an atomic counter, as in adv_tree_kernel):
//   post:  every CTA writes its earliest unfinished body - key and whole record - into row
//          [rank] of the candidate table in global memory
//   barrier A
//   pick:  every CTA reduces the G keys to the same winner, fetches the winner's record,
//          predicts its own bodies, sums its part of the 1 x (N-1) acc+jerk and writes six
//          partial sums into row [rank] of the partial table
//   barrier B
//   owner: adds the G partials in rank order (fixed tree), corrector + new time step.
// Both tables alternate with the step parity, so a fast CTA's post of step s+2 cannot overwrite
// what a slow one still reads in step s (it would have to pass barrier B of step s+1 first).
// Replaces one launch + a mapped-memory round trip per step from the host scheduler (17 us).
constexpr int HGR_THREADS = 256;
constexpr int HGR_PER_CTA = 1600;          // 1600 x 17 doubles = 217.6 KB of the 227 KB
constexpr int HGR_MAX_CTAS = 148;
constexpr int HGR_MAX_N = HGR_MAX_CTAS * HGR_PER_CTA;

struct HermiteGridParams {
  double *state;            // [n][HB]
  double *cand;             // [2][G][HCL_CAND]
  double *part;             // [2][G][8]
  unsigned long long *bar;  // grid barrier counter, zero at launch
  long long *nsteps_out;
  int n;
  double stop_time, eta, eps2;
  long long max_steps;
};

__global__ void __launch_bounds__(HGR_THREADS, 1) hermite_grid_kernel(const HermiteGridParams Q) {
  // [nloc][HRES_STRIDE]; slots 15 / 16 hold the body's rank as two integer keys (hkey_tie of
  // stamp and global index; hkey_of_double of the next time, all ones once finished)
  extern __shared__ double rec[];
  __shared__ double win[HCL_CAND];
  __shared__ double red[HGR_THREADS / 32][6];
  __shared__ unsigned long long loc_key[2];               // warp 0 -> CTA: its earliest body
  __shared__ unsigned long long wkey[HGR_THREADS / 32][2];  // per-warp minima of the G candidates
  constexpr int NW = HGR_THREADS / 32;
  constexpr int NONE = 0x7fffffff;
  const int G = (int)gridDim.x, rank = (int)blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = Q.n;
  const int per = (n + G - 1) / G;
  const int j_lo = min(rank * per, n), j_hi = min(j_lo + per, n), nloc = j_hi - j_lo;
  SweepParams P;
  P.eps2 = Q.eps2;
  for (int l = tid; l < nloc; l += HGR_THREADS) {
#pragma unroll
    for (int c = 0; c < 16; ++c) rec[l * HRES_STRIDE + c] = Q.state[(size_t)(j_lo + l) * HB + c];
  }
  auto K1 = [&](int l) -> unsigned long long & {
    return *reinterpret_cast<unsigned long long *>(rec + l * HRES_STRIDE + 16);
  };
  auto K2 = [&](int l) -> unsigned long long & {
    return *reinterpret_cast<unsigned long long *>(rec + l * HRES_STRIDE + 15);
  };
  for (int l = tid; l < nloc; l += HGR_THREADS) {
    const double *r = rec + l * HRES_STRIDE;
    const double st = r[15];
    K1(l) = hkey_of_double(r[0] + r[14]);
    K2(l) = hkey_tie(st, j_lo + l);
  }
  __syncthreads();
  long long steps = 0, stamp = 0;
  unsigned long long nbar = 0;
  bool finishing = false;
  auto less = [](unsigned long long a1, unsigned long long a2, unsigned long long b1,
                 unsigned long long b2) { return a1 < b1 || (a1 == b1 && a2 < b2); };
  // the CTA's earliest body: cached class minima in warp 0, as in hermite_cluster_kernel
  unsigned long long c1 = HKEY_NONE, c2 = HKEY_NONE;
  if (warp == 0)
    for (int l = lane; l < nloc; l += 32) {
      const unsigned long long k1 = K1(l), k2 = K2(l);
      if (less(k1, k2, c1, c2)) c1 = k1, c2 = k2;
    }
  int changed = -1;  // local index of the body whose key changed in the last step (owner only)
  for (;;) {
    const int par = (int)(steps & 1);
    double *cand = Q.cand + (size_t)par * G * HCL_CAND;
    double *part = Q.part + (size_t)par * G * 8;
    // ---- post: this CTA's earliest body (next time asc, stamp desc, global index asc)
    if (warp == 0) {
      if (changed >= 0) {  // all lanes look at the class of the body that was stepped
        const int g = changed & 31;
        unsigned long long a1 = HKEY_NONE, a2 = HKEY_NONE;
        for (int l = g + 32 * lane; l < nloc; l += 32 * 32) {
          const unsigned long long k1 = K1(l), k2 = K2(l);
          if (less(k1, k2, a1, a2)) a1 = k1, a2 = k2;
        }
        warp_min_pair(a1, a2);
        if (lane == g) c1 = a1, c2 = a2;
      }
      unsigned long long b1 = c1, b2 = c2;
      warp_min_pair(b1, b2);
      if (lane == 0) loc_key[0] = b1, loc_key[1] = b2;
    }
    __syncthreads();
    {
      const unsigned long long k1 = loc_key[0], k2 = loc_key[1];
      const int kidx = k1 == HKEY_NONE ? NONE : (int)(k2 & ((1u << HKEY_IDX_BITS) - 1));
      if (tid < HCL_CAND) {
        double v = 0.0;
        if (tid == 0) v = __longlong_as_double((long long)k1);
        else if (tid == 1) v = __longlong_as_double((long long)k2);
        else if (tid >= 3 && tid < 19 && kidx != NONE) v = rec[(kidx - j_lo) * HRES_STRIDE + (tid - 3)];
        __stcg(cand + (size_t)rank * HCL_CAND + tid, v);
      }
    }
    adv_grid_barrier(Q.bar, (unsigned long long)G * ++nbar);  // A
    // ---- pick: the same winner everywhere (thread r reads CTA r's key; G <= HGR_THREADS)
    unsigned long long w1 = HKEY_NONE, w2 = HKEY_NONE;
    if (tid < G) {
      const double *c = cand + (size_t)tid * HCL_CAND;
      w1 = (unsigned long long)__double_as_longlong(__ldcg(c));
      w2 = (unsigned long long)__double_as_longlong(__ldcg(c + 1));
    }
    warp_min_pair(w1, w2);
    if (lane == 0) wkey[warp][0] = w1, wkey[warp][1] = w2;
    __syncthreads();
    w1 = lane < NW ? wkey[lane][0] : HKEY_NONE;
    w2 = lane < NW ? wkey[lane][1] : HKEY_NONE;
    warp_min_pair(w1, w2);  // every warp ends up with the winner
    if (w1 == HKEY_NONE) break;  // every body finished (uniform over the grid)
    const int it = (int)(w2 & ((1u << HKEY_IDX_BITS) - 1)), owner = it / per;
    if (!finishing) {
      if (steps >= Q.max_steps) break;
      if (hkey_to_double(w1) > Q.stop_time) finishing = true;
    }
    if (tid < HCL_CAND) win[tid] = __ldcg(cand + (size_t)owner * HCL_CAND + tid);
    __syncthreads();
    const double *tr = &win[3];
    const double h = finishing ? __dsub_rn(Q.stop_time, tr[0]) : tr[14];
    const double nt = __dadd_rn(tr[0], h);
    double mt, qt[3], vt[3];
    hermite_predict_fast(tr, nt, mt, qt, vt);
    const OpGradVDGradV::Tgt tg = {qt[0], qt[1], qt[2], vt[0], vt[1], vt[2]};
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int l = tid; l < nloc; l += HGR_THREADS) {
      double mj, qj[3], vj[3];
      hermite_predict_fast(rec + l * HRES_STRIDE, nt, mj, qj, vj);
      const double2 src[4] = {make_double2(qj[0], qj[1]), make_double2(qj[2], mj),
                              make_double2(vj[0], vj[1]), make_double2(vj[2], 0.0)};
      OpGradVDGradV::pair<true>(tg, acc, src, src, (j_lo + l) == it, P);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int c = 0; c < 6; ++c) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], off);
    if (lane == 0)
#pragma unroll
      for (int c = 0; c < 6; ++c) red[warp][c] = acc[c];
    __syncthreads();
    if (tid < 6) {
      double a = red[0][tid];
      for (int w = 1; w < NW; ++w) a += red[w][tid];
      __stcg(part + (size_t)rank * 8 + tid, a);
    }
    adv_grid_barrier(Q.bar, (unsigned long long)G * ++nbar);  // B
    ++steps;
    ++stamp;
    // ---- owner: the G partials in rank order, corrector (hermite.ml:116-123), new time step
    if (rank == owner) {
      if (warp < 6) {  // warp c sums component c: lanes take ranks lane, lane+32, ..., fixed tree
        double a = 0.0;
        for (int r = lane; r < G; r += 32) a += __ldcg(part + (size_t)r * 8 + warp);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
        if (lane == 0) red[0][warp] = a;
      }
      __syncthreads();
      if (tid == 0) {
        double fp[3], dfp[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          fp[c] = mt * red[0][c];
          dfp[c] = mt * red[0][3 + c];
        }
        double *wr = rec + (it - j_lo) * HRES_STRIDE;
        const double m = mt, eta = Q.eta;
        double qn[3], pn[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double q0 = tr[2 + c], p0 = tr[5 + c], f0 = tr[8 + c], df0 = tr[11 + c];
          pn[c] = __dadd_rn(__dadd_rn(p0, __dmul_rn(__dmul_rn(0.5, h), __dadd_rn(f0, fp[c]))),
                            __dmul_rn(__ddiv_rn(__dmul_rn(h, h), 12.0), __dsub_rn(df0, dfp[c])));
          qn[c] = __dadd_rn(
              __dadd_rn(q0, __dmul_rn(__ddiv_rn(__dmul_rn(0.5, h), m), __dadd_rn(p0, pn[c]))),
              __dmul_rn(__ddiv_rn(__dmul_rn(h, h), __dmul_rn(12.0, m)), __dsub_rn(f0, fp[c])));
        }
        const double dx = qn[0] - qt[0], dy = qn[1] - qt[1], dz = qn[2] - qt[2];
        const double rr = __dsqrt_rn(
            __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        const double fn = __dsqrt_rn(__dadd_rn(
            __dadd_rn(__dmul_rn(fp[0], fp[0]), __dmul_rn(fp[1], fp[1])), __dmul_rn(fp[2], fp[2])));
        double hn;
        if (rr == 0.0) {
          hn = __dmul_rn(h, 1.189207115002721);
        } else {
          const double x =
              __ddiv_rn(__dmul_rn(__dmul_rn(eta, __dmul_rn(h, h)), fn), __dmul_rn(m, rr));
          hn = __dmul_rn(h, __dsqrt_rn(__dsqrt_rn(x)));
        }
        wr[0] = nt;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          wr[2 + c] = qn[c];
          wr[5 + c] = pn[c];
          wr[8 + c] = fp[c];
          wr[11 + c] = dfp[c];
        }
        wr[14] = hn;
        K1(it - j_lo) = finishing ? HKEY_NONE : hkey_of_double(nt + hn);
        K2(it - j_lo) = hkey_tie((double)stamp, it);
      }
    }
    changed = rank == owner ? it - j_lo : -1;
    __syncthreads();  // the owner's next search must see the update; uniform in every CTA
  }
  __syncthreads();
  for (int l = tid; l < nloc; l += HGR_THREADS) {
    double *rl = rec + l * HRES_STRIDE;
    rl[15] = (double)(0x7ffffffffffull - (K2(l) >> HKEY_IDX_BITS));  // back to the stamp
#pragma unroll
    for (int c = 0; c < 16; ++c) Q.state[(size_t)(j_lo + l) * HB + c] = rl[c];
  }
  if (rank == 0 && tid == 0) *Q.nsteps_out = steps;
}

// ---- the whole-grid loop, software-pipelined ---------------------------------------------------
// The scheme of hermite_cluster_spec_kernel with the tables in global memory (L2) and the grid
// barrier at the end of the step - ONE per step instead of two, and the owner's corrector runs
// beside the other CTAs' work:
//   thread 0 of i_k's owner   corrector of step k (the G partials are first added by six warps of
//                             the owner, one component each, in a fixed tree), new keys and
//                             record -> row 2 G of the candidate table
//   warp 1 of every CTA       the CTA's two earliest bodies other than i_k -> rows 2 rank, 2 rank + 1
//   warps 2.. of every CTA    local part of the next target's sums without i_k -> row [rank] of the
//                             partial table
//   grid barrier
//   everybody                 the two smallest keys of the 2 G + 1 rows: next target and the one
//                             after; their records (and the stepped body's) come into shared memory.
constexpr int HGRS_THREADS = 320;

struct HermiteGridSpecParams {
  double *state;            // [n][HB]
  double *cand;             // [2][2 G + 1][HCL_CAND]
  double *part;             // [2][G][8]
  unsigned long long *bar;  // grid barrier counter, zero at launch
  long long *nsteps_out;
  int n;
  double stop_time, eta, eps2;
  long long max_steps;
};

__global__ void __launch_bounds__(HGRS_THREADS, 1)
    hermite_grid_spec_kernel(const HermiteGridSpecParams Q) {
  constexpr int T = HGRS_THREADS, NW = T / 32, NF = NW - 2;
  constexpr int NONE = 0x7fffffff;
  extern __shared__ double rec[];  // [nloc][HRES_STRIDE]; slots 15 / 16 = the integer keys
  __shared__ double rows3[2][3][HCL_CAND];  // by step parity: current target, next target, stepped body
  __shared__ double tgtp[2][8];             // the next target's predicted state (local copy)
  __shared__ double red[NW][6];
  __shared__ double psum[6];                // the owner: the G partials added up
  __shared__ unsigned long long wtop[NW][2][2];
  __shared__ int wrow[NW][2];
  const int G = (int)gridDim.x, rank = (int)blockIdx.x;
  const int NCAND = 2 * G + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = Q.n;
  const int per = (n + G - 1) / G;
  const int j_lo = min(rank * per, n), j_hi = min(j_lo + per, n), nloc = j_hi - j_lo;
  const double stop_time = Q.stop_time, eta = Q.eta;
  const long long max_steps = Q.max_steps;
  SweepParams P;
  P.eps2 = Q.eps2;
  auto K1 = [&](int l) -> unsigned long long & {
    return *reinterpret_cast<unsigned long long *>(rec + l * HRES_STRIDE + 16);
  };
  auto K2 = [&](int l) -> unsigned long long & {
    return *reinterpret_cast<unsigned long long *>(rec + l * HRES_STRIDE + 15);
  };
  for (int l = tid; l < nloc; l += T) {
    double *r = rec + l * HRES_STRIDE;
#pragma unroll
    for (int c = 0; c < 16; ++c) r[c] = Q.state[(size_t)(j_lo + l) * HB + c];
    const double st = r[15];
    K1(l) = hkey_of_double(r[0] + r[14]);
    K2(l) = hkey_tie(st, j_lo + l);
  }
  __syncthreads();
  auto less = [](unsigned long long a1, unsigned long long a2, unsigned long long b1,
                 unsigned long long b2) { return a1 < b1 || (a1 == b1 && a2 < b2); };
  auto key_of = [](const double *c, int w) {
    return (unsigned long long)__double_as_longlong(c[w]);
  };
  auto idx_of = [](unsigned long long k1, unsigned long long k2) {
    return k1 == HKEY_NONE ? NONE : (int)(k2 & ((1u << HKEY_IDX_BITS) - 1));
  };
  auto keep2 = [&](unsigned long long k1, unsigned long long k2, unsigned long long &a1,
                   unsigned long long &a2, unsigned long long &b1, unsigned long long &b2) {
    const bool lt_a = less(k1, k2, a1, a2), lt_b = less(k1, k2, b1, b2);
    b1 = lt_a ? a1 : (lt_b ? k1 : b1);
    b2 = lt_a ? a2 : (lt_b ? k2 : b2);
    a1 = lt_a ? k1 : a1;
    a2 = lt_a ? k2 : a2;
  };
  auto combine2 = [&](unsigned long long a1, unsigned long long a2, unsigned long long b1,
                      unsigned long long b2, unsigned long long (&o)[2][2]) {
    unsigned long long g1 = a1, g2 = a2;
    warp_min_pair(g1, g2);
    const bool mine = (a1 == g1 && a2 == g2);  // keys are unique (the tie key holds the index)
    unsigned long long s1 = mine ? b1 : a1, s2 = mine ? b2 : a2;
    warp_min_pair(s1, s2);
    o[0][0] = g1, o[0][1] = g2, o[1][0] = s1, o[1][1] = s2;
  };
  // warp 1: this CTA's two earliest bodies other than `hide` (search as in
  // hermite_cluster_spec_kernel) -> rows 2 rank, 2 rank + 1 of cand[pb]
  const bool cached = nloc > 384;
  unsigned long long ca1 = HKEY_NONE, ca2 = HKEY_NONE, cb1 = HKEY_NONE, cb2 = HKEY_NONE;
  auto rescan = [&](int g, int hide) {
    unsigned long long a1 = HKEY_NONE, a2 = HKEY_NONE, b1 = HKEY_NONE, b2 = HKEY_NONE;
    for (int l = g + 32 * lane; l < nloc; l += 32 * 32)
      keep2(l == hide ? HKEY_NONE : K1(l), K2(l), a1, a2, b1, b2);
    unsigned long long o[2][2];
    combine2(a1, a2, b1, b2, o);
    if (lane == g) ca1 = o[0][0], ca2 = o[0][1], cb1 = o[1][0], cb2 = o[1][1];
  };
  auto post_candidates = [&](int pb, int hide, int changed) {
    unsigned long long a1 = HKEY_NONE, a2 = HKEY_NONE, b1 = HKEY_NONE, b2 = HKEY_NONE;
    if (cached) {
      if (changed >= 0) rescan(changed & 31, hide);
      if (hide >= 0 && (changed < 0 || (changed & 31) != (hide & 31))) rescan(hide & 31, hide);
      a1 = ca1, a2 = ca2, b1 = cb1, b2 = cb2;
    } else {
#pragma unroll 4
      for (int l = lane; l < nloc; l += 32) keep2(l == hide ? HKEY_NONE : K1(l), K2(l), a1, a2, b1, b2);
    }
    unsigned long long o[2][2];
    combine2(a1, a2, b1, b2, o);
    for (int v = lane; v < 2 * HCL_CAND; v += 32) {
      const int s = v / HCL_CAND, slot = v - s * HCL_CAND;
      const int kl = o[s][0] == HKEY_NONE ? -1 : (int)(o[s][1] & ((1u << HKEY_IDX_BITS) - 1)) - j_lo;
      double x = 0.0;
      if (slot == 0) x = __longlong_as_double((long long)o[s][0]);
      else if (slot == 1) x = __longlong_as_double((long long)o[s][1]);
      else if (slot >= 3 && slot < 19 && kl >= 0) x = rec[kl * HRES_STRIDE + (slot - 3)];
      __stcg(Q.cand + ((size_t)pb * NCAND + 2 * rank + s) * HCL_CAND + slot, x);
    }
  };
  // everybody: the two smallest keys of the first nrows rows of cand[pb] -> row numbers (-1: none),
  // then those rows and row 2 G come into rows3[pb] (one block barrier inside, one at the end)
  auto pick2 = [&](int pb, int nrows, int &row1, int &row2) {
    const double *tab = Q.cand + (size_t)pb * NCAND * HCL_CAND;
    unsigned long long a1 = HKEY_NONE, a2 = HKEY_NONE, b1 = HKEY_NONE, b2 = HKEY_NONE;
    int ra = -1, rb = -1;
    for (int x = tid; x < nrows; x += T) {  // G <= 148: one row per thread
      const unsigned long long k1 = (unsigned long long)__double_as_longlong(__ldcg(tab + (size_t)x * HCL_CAND)),
                               k2 = (unsigned long long)__double_as_longlong(__ldcg(tab + (size_t)x * HCL_CAND + 1));
      if (less(k1, k2, a1, a2)) b1 = a1, b2 = a2, rb = ra, a1 = k1, a2 = k2, ra = x;
      else if (less(k1, k2, b1, b2)) b1 = k1, b2 = k2, rb = x;
    }
    // warp top-2 with the rows they sit in
    unsigned long long g1 = a1, g2 = a2;
    warp_min_pair(g1, g2);
    const bool mine = a1 == g1 && a2 == g2 && g1 != HKEY_NONE;
    const unsigned m1 = __ballot_sync(0xffffffffu, mine);
    const int src1 = m1 ? __ffs(m1) - 1 : 0;
    const int r1 = __shfl_sync(0xffffffffu, ra, src1);
    unsigned long long s1 = mine && lane == src1 ? b1 : a1, s2 = mine && lane == src1 ? b2 : a2;
    const int rs = mine && lane == src1 ? rb : ra;
    unsigned long long h1 = s1, h2 = s2;
    warp_min_pair(h1, h2);
    const bool sec = s1 == h1 && s2 == h2 && h1 != HKEY_NONE;
    const unsigned m2 = __ballot_sync(0xffffffffu, sec);
    const int src2 = m2 ? __ffs(m2) - 1 : 0;
    const int r2 = __shfl_sync(0xffffffffu, rs, src2);
    if (lane == 0) {
      wtop[warp][0][0] = g1, wtop[warp][0][1] = g2, wtop[warp][1][0] = h1, wtop[warp][1][1] = h2;
      wrow[warp][0] = m1 ? r1 : -1;
      wrow[warp][1] = m2 ? r2 : -1;
    }
    __syncthreads();
    // every warp: the two smallest of the 2 NW warp results (lane x: warp x / 2, entry x & 1)
    a1 = a2 = HKEY_NONE;
    ra = -1;
    if (lane < 2 * NW) {
      a1 = wtop[lane >> 1][lane & 1][0];
      a2 = wtop[lane >> 1][lane & 1][1];
      ra = wrow[lane >> 1][lane & 1];
    }
    g1 = a1, g2 = a2;
    warp_min_pair(g1, g2);
    const bool mine_b = a1 == g1 && a2 == g2 && g1 != HKEY_NONE;
    const unsigned n1 = __ballot_sync(0xffffffffu, mine_b);
    const int t1 = n1 ? __ffs(n1) - 1 : 0;
    row1 = n1 ? __shfl_sync(0xffffffffu, ra, t1) : -1;
    s1 = mine_b && lane == t1 ? HKEY_NONE : a1;
    s2 = mine_b && lane == t1 ? HKEY_NONE : a2;
    h1 = s1, h2 = s2;
    warp_min_pair(h1, h2);
    const bool sec_b = s1 == h1 && s2 == h2 && h1 != HKEY_NONE;
    const unsigned n2 = __ballot_sync(0xffffffffu, sec_b);
    const int t2 = n2 ? __ffs(n2) - 1 : 0;
    row2 = n2 ? __shfl_sync(0xffffffffu, ra, t2) : -1;
    // the three records into shared memory
    if (tid < 3 * HCL_CAND) {
      const int which = tid / HCL_CAND, slot = tid - which * HCL_CAND;
      const int row = which == 0 ? row1 : (which == 1 ? row2 : 2 * G);
      rows3[pb][which][slot] = row >= 0 && row < nrows ? __ldcg(tab + (size_t)row * HCL_CAND + slot) : 0.0;
    }
    __syncthreads();
  };
  // warps [w0, NW): the local part of the six sums -> row [rank] of part[pb]
  auto local_force = [&](int w0, int hide, int itarget, const OpGradVDGradV::Tgt &tg, double nt,
                         int pb, auto sync) {
    const int nthr = (NW - w0) * 32, t0 = tid - w0 * 32;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int l = t0; l < nloc; l += nthr) {
      if (l == hide) continue;
      double mj, qj[3], vj[3];
      hermite_predict_fast(rec + l * HRES_STRIDE, nt, mj, qj, vj);
      const double2 src[4] = {make_double2(qj[0], qj[1]), make_double2(qj[2], mj),
                              make_double2(vj[0], vj[1]), make_double2(vj[2], 0.0)};
      OpGradVDGradV::pair<true>(tg, acc, src, src, (j_lo + l) == itarget, P);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int c = 0; c < 6; ++c) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], off);
    if (lane == 0)
#pragma unroll
      for (int c = 0; c < 6; ++c) red[warp][c] = acc[c];
    sync();
    if (t0 < 6) {
      double a = red[w0][t0];
      for (int w = w0 + 1; w < NW; ++w) a += red[w][t0];
      __stcg(Q.part + ((size_t)pb * G + rank) * 8 + t0, a);
    }
  };
  auto sync_all = [] { __syncthreads(); };
  auto sync_force = [] { asm volatile("bar.sync 2, %0;" ::"r"(NF * 32) : "memory"); };
  // owner: warps 2..7 add the G partials of part[pb], one component each, fixed tree -> psum;
  // hands over to warp 0 through a named barrier (warps 0 and 2..7: 224 threads)
  auto collect = [&](int pb) {
    if (warp >= 2 && warp < 8) {
      const int c = warp - 2;
      double a = 0.0;
      for (int r = lane; r < G; r += 32) a += __ldcg(Q.part + ((size_t)pb * G + r) * 8 + c);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
      if (lane == 0) psum[c] = a;
    }
    if (warp == 0 || (warp >= 2 && warp < 8)) asm volatile("bar.sync 3, 224;" ::: "memory");
  };

  long long steps = 0, stamp = 0;
  unsigned long long nbar = 0;
  bool finishing = false;
  // ---- the first two bodies of the schedule
  if (warp == 1) {
    if (cached) {
#pragma unroll 4
      for (int l = lane; l < nloc; l += 32) keep2(K1(l), K2(l), ca1, ca2, cb1, cb2);
    }
    post_candidates(1, -1, -1);
  }
  adv_grid_barrier(Q.bar, (unsigned long long)G * ++nbar);
  int row1, row2;
  pick2(1, 2 * G, row1, row2);
  const double *cur = row1 >= 0 ? &rows3[1][0][0] : nullptr;  // the current target: keys, -, record
  const double *nxt = row2 >= 0 ? &rows3[1][1][0] : nullptr;  // the target after it
  int par = 0, pb = 0;
  bool have = false;
  const double *prevrec = nullptr;  // the body stepped before (its new record), for the extra pair
  int changed = -1;
  for (;;) {
    if (cur == nullptr) break;  // every body finished (uniform over the grid)
    const unsigned long long itk1 = key_of(cur, 0);
    const int it = idx_of(itk1, key_of(cur, 1)), owner = it / per;
    if (!finishing) {
      if (steps >= max_steps) break;
      if (hkey_to_double(itk1) > stop_time) finishing = true;
    }
    const double *tr = cur + 3;
    const double h = finishing ? __dsub_rn(stop_time, tr[0]) : tr[14];
    const double nt = __dadd_rn(tr[0], h);
    const int it_loc = rank == owner ? it - j_lo : -1;
    double mt = 0.0, qt[3] = {0, 0, 0}, vt[3] = {0, 0, 0};
    if (!have) {
      // the plain way: everybody sums its local part for `it`, one more grid barrier
      hermite_predict_fast(tr, nt, mt, qt, vt);
      const OpGradVDGradV::Tgt tg = {qt[0], qt[1], qt[2], vt[0], vt[1], vt[2]};
      local_force(0, -1, it, tg, nt, pb, sync_all);
      adv_grid_barrier(Q.bar, (unsigned long long)G * ++nbar);
    } else if (tid == 0 && rank == owner) {
      mt = tgtp[pb][0];
#pragma unroll
      for (int c = 0; c < 3; ++c) qt[c] = tgtp[pb][1 + c], vt[c] = tgtp[pb][4 + c];
    }
    ++steps;
    ++stamp;
    const int r = nxt ? idx_of(key_of(nxt, 0), key_of(nxt, 1)) : NONE;
    if (rank == owner) collect(pb);
    if (warp == 0) {
      if (tid == 0 && rank == owner) {
        // ---- corrector (hermite.ml:116-123), new time step (hermite.ml:87-93)
        const OpGradVDGradV::Tgt tg = {qt[0], qt[1], qt[2], vt[0], vt[1], vt[2]};
        double extra[6] = {0, 0, 0, 0, 0, 0};
        if (have) {  // the interaction the speculative sums left out: the body stepped last
          double mj, qj[3], vj[3];
          hermite_predict_fast(prevrec + 3, nt, mj, qj, vj);
          const double2 src[4] = {make_double2(qj[0], qj[1]), make_double2(qj[2], mj),
                                  make_double2(vj[0], vj[1]), make_double2(vj[2], 0.0)};
          OpGradVDGradV::pair<false>(tg, extra, src, src, false, P);
        }
        double fp[3], dfp[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          fp[c] = mt * (extra[c] + psum[c]);
          dfp[c] = mt * (extra[3 + c] + psum[3 + c]);
        }
        double *wr = rec + it_loc * HRES_STRIDE;
        const double m = mt;
        double qn[3], pn[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double q0 = tr[2 + c], p0 = tr[5 + c], f0 = tr[8 + c], df0 = tr[11 + c];
          pn[c] = __dadd_rn(__dadd_rn(p0, __dmul_rn(__dmul_rn(0.5, h), __dadd_rn(f0, fp[c]))),
                            __dmul_rn(__ddiv_rn(__dmul_rn(h, h), 12.0), __dsub_rn(df0, dfp[c])));
          qn[c] = __dadd_rn(
              __dadd_rn(q0, __dmul_rn(__ddiv_rn(__dmul_rn(0.5, h), m), __dadd_rn(p0, pn[c]))),
              __dmul_rn(__ddiv_rn(__dmul_rn(h, h), __dmul_rn(12.0, m)), __dsub_rn(f0, fp[c])));
        }
        const double dx = qn[0] - qt[0], dy = qn[1] - qt[1], dz = qn[2] - qt[2];
        const double rr = __dsqrt_rn(
            __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        const double fn = __dsqrt_rn(__dadd_rn(
            __dadd_rn(__dmul_rn(fp[0], fp[0]), __dmul_rn(fp[1], fp[1])), __dmul_rn(fp[2], fp[2])));
        double hn;
        if (rr == 0.0) {
          hn = __dmul_rn(h, 1.189207115002721);
        } else {
          const double x =
              __ddiv_rn(__dmul_rn(__dmul_rn(eta, __dmul_rn(h, h)), fn), __dmul_rn(m, rr));
          hn = __dmul_rn(h, __dsqrt_rn(__dsqrt_rn(x)));
        }
        wr[0] = nt;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          wr[2 + c] = qn[c];
          wr[5 + c] = pn[c];
          wr[8 + c] = fp[c];
          wr[11 + c] = dfp[c];
        }
        wr[14] = hn;
        K1(it_loc) = finishing ? HKEY_NONE : hkey_of_double(nt + hn);
        K2(it_loc) = hkey_tie((double)stamp, it);
      }
      __syncwarp();
      if (rank == owner && lane < 19) {  // the new keys and record -> row 2 G
        const double *wr = rec + it_loc * HRES_STRIDE;
        double v = 0.0;
        if (lane == 0) v = wr[16];
        else if (lane == 1) v = wr[15];
        else if (lane >= 3) v = wr[lane - 3];
        __stcg(Q.cand + ((size_t)par * NCAND + 2 * G) * HCL_CAND + lane, v);
      }
    } else if (warp == 1) {
      post_candidates(par, it_loc, changed);
    } else if (r != NONE) {
      const double *rr_ = nxt + 3;
      const bool fin_r = finishing || hkey_to_double(key_of(nxt, 0)) > stop_time;
      const double hr = fin_r ? __dsub_rn(stop_time, rr_[0]) : rr_[14];
      const double ntr = __dadd_rn(rr_[0], hr);
      double mr, qr[3], vr[3];
      hermite_predict_fast(rr_, ntr, mr, qr, vr);
      const OpGradVDGradV::Tgt tgr = {qr[0], qr[1], qr[2], vr[0], vr[1], vr[2]};
      if (tid == 64) {
        tgtp[pb ^ 1][0] = mr;
#pragma unroll
        for (int c = 0; c < 3; ++c) tgtp[pb ^ 1][1 + c] = qr[c], tgtp[pb ^ 1][4 + c] = vr[c];
      }
      local_force(2, it_loc, r, tgr, ntr, pb ^ 1, sync_force);
    }
    adv_grid_barrier(Q.bar, (unsigned long long)G * ++nbar);
    // ---- who is next, and who after that (uniform over the grid)
    pick2(par, NCAND, row1, row2);
    changed = it_loc;
    const bool again = row1 == 2 * G;
    prevrec = &rows3[par][2][0];
    cur = row1 >= 0 ? &rows3[par][0][0] : nullptr;
    nxt = row2 >= 0 ? &rows3[par][1][0] : nullptr;
    have = !again && cur != nullptr && idx_of(key_of(cur, 0), key_of(cur, 1)) == r;
    if (have) pb ^= 1;
    par ^= 1;
  }
  __syncthreads();
  for (int l = tid; l < nloc; l += T) {
    double *rl = rec + l * HRES_STRIDE;
    rl[15] = (double)(0x7ffffffffffull - (K2(l) >> HKEY_IDX_BITS));  // back to the stamp
#pragma unroll
    for (int c = 0; c < 16; ++c) Q.state[(size_t)(j_lo + l) * HB + c] = rl[c];
  }
  if (rank == 0 && tid == 0) *Q.nsteps_out = steps;
}

// ---- Analysis.binaries support ------------------------------------------------------------------
// Slot 7 of a packed body := its global index (OpBinary reports partners by index).
__global__ void index_kernel(double *__restrict__ packed, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) packed[(size_t)i * PK + 7] = (double)i;
}

// Lists the pairs (i, j > i) with binary_energy < thr for the flagged targets (those whose
// scan count is non-zero; there are few): one CTA per flagged target, threads stride over j.
// Pairs are appended in arrival order; the caller sorts them.  Same IEEE sequence as OpBinary.
__global__ void __launch_bounds__(256) binary_list_kernel(const double *__restrict__ packed, int n,
                                                          const int *__restrict__ flagged,
                                                          double eps2, double thr, int max_pairs,
                                                          int *__restrict__ pi, int *__restrict__ pj,
                                                          double *__restrict__ pe,
                                                          int *__restrict__ npairs) {
  const int i = flagged[blockIdx.x];
  OpBinary::Tgt t;
  SweepParams P;
  P.eps2 = eps2;
  P.dt = thr;
  OpBinary::load(t, packed + (size_t)i * PK, P, i);
  for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
    double acc[3];
    OpBinary::zero(acc);
    OpBinary::pair<true>(t, acc, reinterpret_cast<const double2 *>(packed + (size_t)j * PK),
                         nullptr, false, P);
    if (acc[2] > 0.0) {
      const int slot = atomicAdd(npairs, 1);
      if (slot < max_pairs) {
        pi[slot] = i;
        pj[slot] = j;
        pe[slot] = acc[0];
      }
    }
  }
}

// ---- Leapfrog DKD on device-resident packed bodies (leapfrog.ml:64-68, 129-137) ----------------
// drift: q <- q + dt v with the reference's separate multiply and add.
__global__ void leapfrog_drift_kernel(double *__restrict__ packed, int n, double dt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double *b = packed + (size_t)i * PK;
  b[0] = __dadd_rn(b[0], __dmul_rn(dt, b[4]));
  b[1] = __dadd_rn(b[1], __dmul_rn(dt, b[5]));
  b[2] = __dadd_rn(b[2], __dmul_rn(dt, b[6]));
}
// kick: v <- v + dv (dv = dt * sum_j m_j d / r^3 from the kick sweep).
__global__ void leapfrog_apply_kick_kernel(double *__restrict__ packed,
                                           const double *__restrict__ dv, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double *b = packed + (size_t)i * PK;
  b[4] = __dadd_rn(b[4], dv[3 * i]);
  b[5] = __dadd_rn(b[5], dv[3 * i + 1]);
  b[6] = __dadd_rn(b[6], dv[3 * i + 2]);
}
__global__ void unpack_kernel(const double *__restrict__ packed, int n, double *__restrict__ q,
                              double *__restrict__ v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *b = packed + (size_t)i * PK;
  for (int k = 0; k < 3; ++k) {
    q[3 * i + k] = b[k];
    v[3 * i + k] = b[4 + k];
  }
}

// ---- Advancer.A on device-resident records (advancer.ml:143-252) -------------------------------
// The Advancer.A.body records (advancer.ml:27-45) stay on the device as [n][35] doubles
//   0 t, 1 m, 2-4 q, 5-7 p, 8 h, 9 hmax, 10 t0, 11-13 dVdq0, 14-16 dVdq1, 17-19 dVdq2,
//   20-22 p0, 23-25 q0, 26-28 q1, 29-31 q2, 32-34 cs
// in the reference's array order; the host keeps only the control flow (block recursion, hmax,
// sort permutations).  Every formula repeats the reference's operation sequence with IEEE
// intrinsics (no contraction), so this path is bit-identical to the host mirror.
constexpr int AB = 35;

// Row operations, shared by the per-phase kernels below and by the device-side block recursion
// (adv_tree_kernel).  Loads go through L2 (ld.global.cg): inside the persistent kernel a row is
// written and read by different CTAs across grid barriers.
// A row is reached through an accessor: AdvRowAoS = record i of the [n][35] array (the layout of
// the ABI and of the per-phase kernels); AdvRowSoA = column-major copy the persistent kernel
// works on (field k of row i at base[k*ld + i]: a warp's loads of one field coalesce).
struct AdvRowAoS {
  double *b;
  __device__ __forceinline__ double ld(int k) const { return __ldcg(b + k); }
  __device__ __forceinline__ void st(int k, double v) const { b[k] = v; }
};
struct AdvRowSoA {
  double *b;
  size_t fs;
  __device__ __forceinline__ double ld(int k) const { return __ldcg(b + (size_t)k * fs); }
  __device__ __forceinline__ void st(int k, double v) const { b[(size_t)k * fs] = v; }
};

// predict_q012 b hnew; clear_before_step b  (advancer.ml:143-171)
template <class Row>
__device__ __forceinline__ void adv_begin_row(const Row b, double hnew) {
  const double h = b.ld(8), m = b.ld(1);
  const double hnew2 = __dmul_rn(hnew, hnew), h2 = __dmul_rn(h, h);
  const double hnew3 = __dmul_rn(hnew2, hnew), h3 = __dmul_rn(h2, h);
  const double a0 = __ddiv_rn(__dmul_rn(hnew3, __dadd_rn(h, hnew)), __dmul_rn(__dmul_rn(8.0, h3), m));
  const double a1 = __ddiv_rn(__dmul_rn(hnew3, __dadd_rn(__dmul_rn(2.0, h), hnew)),
                              __dmul_rn(__dmul_rn(16.0, h3), m));
  const double a2 = __ddiv_rn(
      __dmul_rn(hnew2, __dadd_rn(__dadd_rn(__dmul_rn(6.0, h2), __dmul_rn(__dmul_rn(3.0, h), hnew)), hnew2)),
      __dmul_rn(__dmul_rn(8.0, h3), m));
  const double b0 = __ddiv_rn(__dmul_rn(hnew3, __dadd_rn(h, hnew)), __dmul_rn(h3, m));
  const double b1 = __ddiv_rn(__dmul_rn(hnew3, __dadd_rn(__dmul_rn(2.0, h), hnew)),
                              __dmul_rn(__dmul_rn(2.0, h3), m));
  const double b2 = __ddiv_rn(
      __dmul_rn(hnew2, __dadd_rn(__dadd_rn(__dmul_rn(3.0, h2), __dmul_rn(__dmul_rn(3.0, h), hnew)), hnew2)),
      __dmul_rn(h3, m));
  for (int k = 0; k < 3; ++k) {
    const double p = b.ld(5 + k), q = b.ld(2 + k);
    const double v0 = b.ld(11 + k), v1 = b.ld(14 + k), v2 = b.ld(17 + k);
    b.st(20 + k, p);
    b.st(23 + k, q);
    b.st(26 + k, __dsub_rn(
        __dadd_rn(__dsub_rn(__dadd_rn(q, __ddiv_rn(__dmul_rn(hnew, p), __dmul_rn(2.0, m))),
                            __dmul_rn(a0, v0)),
                  __dmul_rn(a1, v1)),
        __dmul_rn(a2, v2)));
    b.st(29 + k, __dsub_rn(
        __dadd_rn(__dsub_rn(__dadd_rn(q, __ddiv_rn(__dmul_rn(hnew, p), m)), __dmul_rn(b0, v0)),
                  __dmul_rn(b1, v1)),
        __dmul_rn(b2, v2)));
  }
  b.st(8, hnew);
  b.st(10, b.ld(0));
  for (int k = 0; k < 3; ++k) {
    b.st(11 + k, 0.0);
    b.st(14 + k, 0.0);
    b.st(17 + k, 0.0);
    b.st(32 + k, 0.0);
  }
  b.st(32, 1.0);
}

// predict_position (advancer.ml:173-195) on values: weights c[3] and q = c0 q0 + c1 q1 + c2 q2
__device__ __forceinline__ void adv_predict_vals(double t, double t0, double h, const double *q0,
                                                 const double *q1, const double *q2, double *c,
                                                 double *q) {
  const double dt = __dsub_rn(t, t0);
  const double h2 = __dmul_rn(h, h);
  c[0] = __ddiv_rn(
      __dadd_rn(__dsub_rn(__dmul_rn(__dmul_rn(2.0, dt), dt), __dmul_rn(__dmul_rn(3.0, dt), h)), h2), h2);
  c[1] = __ddiv_rn(__dmul_rn(__dmul_rn(4.0, dt), __dsub_rn(h, dt)), h2);
  c[2] = __ddiv_rn(__dmul_rn(dt, __dsub_rn(__dmul_rn(2.0, dt), h)), h2);
  for (int k = 0; k < 3; ++k)
    q[k] = __dadd_rn(__dadd_rn(__dmul_rn(c[0], q0[k]), __dmul_rn(c[1], q1[k])), __dmul_rn(c[2], q2[k]));
}

// predict_position b t; returns {q, m} as the two double2 of a packed row and the row's cs
template <class Row>
__device__ __forceinline__ void adv_predict_row(const Row b, double t, double2 &xy, double2 &zm,
                                                double *cs) {
  // every load the usual branch needs is issued up front: one trip to L2 instead of three
  const double told = b.ld(0), t0 = b.ld(10), h = b.ld(8), m = b.ld(1);
  double q0[3], q1[3], q2[3];
  for (int k = 0; k < 3; ++k) {
    q0[k] = b.ld(23 + k);
    q1[k] = b.ld(26 + k);
    q2[k] = b.ld(29 + k);
  }
  if (told != t) {
    double q[3];
    adv_predict_vals(t, t0, h, q0, q1, q2, cs, q);
    for (int k = 0; k < 3; ++k) {
      b.st(32 + k, cs[k]);
      b.st(2 + k, q[k]);
    }
    b.st(0, t);
    xy = make_double2(q[0], q[1]);
    zm = make_double2(q[2], m);
  } else {
    for (int k = 0; k < 3; ++k) cs[k] = b.ld(32 + k);
    xy = make_double2(b.ld(2), b.ld(3));
    zm = make_double2(b.ld(4), m);
  }
}

// dVdqK += cK S, cK = vC cs[K]  (advancer.ml:216-234 in per-target form)
template <class Row>
__device__ __forceinline__ void adv_accumulate_row(const Row b, double c0, double c1, double c2,
                                                   const double *s) {
  for (int k = 0; k < 3; ++k) {
    b.st(11 + k, __dadd_rn(b.ld(11 + k), __dmul_rn(c0, s[k])));
    b.st(14 + k, __dadd_rn(b.ld(14 + k), __dmul_rn(c1, s[k])));
    b.st(17 + k, __dadd_rn(b.ld(17 + k), __dmul_rn(c2, s[k])));
  }
}

// the same with the nine accumulators already in registers (loaded ahead of the pair loop)
template <class Row>
__device__ __forceinline__ void adv_accumulate_row_pre(const Row b, const double *v, double c0,
                                                       double c1, double c2, const double *s) {
  for (int k = 0; k < 3; ++k) {
    b.st(11 + k, __dadd_rn(v[k], __dmul_rn(c0, s[k])));
    b.st(14 + k, __dadd_rn(v[3 + k], __dmul_rn(c1, s[k])));
    b.st(17 + k, __dadd_rn(v[6 + k], __dmul_rn(c2, s[k])));
  }
}

// finish_body (advancer.ml:238-252) on values: v = dVdq0..2 (9), p0, q0 -> q, p
__device__ __forceinline__ void adv_finish_vals(double h, double m, const double *v, const double *p0,
                                                const double *q0, double *q, double *p) {
  const double cp = __ddiv_rn(h, m);
  const double cV0 = cp, cV1 = __dmul_rn(0.5, cp);
  for (int k = 0; k < 3; ++k) {
    const double v0 = v[k], v1 = v[3 + k], v2 = v[6 + k], pi = p0[k];
    q[k] = __dsub_rn(__dsub_rn(__dadd_rn(q0[k], __dmul_rn(cp, pi)), __dmul_rn(cV0, v0)),
                     __dmul_rn(cV1, v1));
    p[k] = __dsub_rn(__dsub_rn(__dsub_rn(pi, v0), v1), v2);
  }
}

// finish_body b t
template <class Row>
__device__ __forceinline__ void adv_finish_row(const Row b, double t) {
  const double h = b.ld(8), m = b.ld(1);
  double v[9], p0[3], q0[3], q[3], p[3];
  for (int k = 0; k < 9; ++k) v[k] = b.ld(11 + k);
  for (int k = 0; k < 3; ++k) {
    p0[k] = b.ld(20 + k);
    q0[k] = b.ld(23 + k);
  }
  adv_finish_vals(h, m, v, p0, q0, q, p);
  for (int k = 0; k < 3; ++k) {
    b.st(2 + k, q[k]);
    b.st(5 + k, p[k]);
  }
  b.st(0, t);
}

// h_adaptive sf b (advancer.ml:257-281) -> the body's new hmax, on values (v2 = dVdq2).  The
// operation sequence is the reference's; `**` is CUDA's pow (<= 2 ulp), where the host paths use
// libm's: hmax only steers comparisons (block membership, sort order), so a last-bit difference
// changes nothing unless it flips one of them.
__device__ __forceinline__ double adv_h_adaptive_vals(double h, double m, double t, const double *q,
                                                      const double *q2, const double *v2, double sf) {
  const double dx = __dsub_rn(q[0], q2[0]), dy = __dsub_rn(q[1], q2[1]), dz = __dsub_rn(q[2], q2[2]);
  const double dq =
      __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
  double dt;
  if (dq == 0.0) {
    dt = __dmul_rn(h, 2.01);
  } else {
    const double vx = v2[0], vy = v2[1], vz = v2[2];
    const double nv =
        __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
    const double a = __ddiv_rn(__ddiv_rn(__dmul_rn(nv, 6.0), h), m);
    const double a_dq = __dmul_rn(__dmul_rn(__dmul_rn(0.5, h), h), a);
    const double ratio = __ddiv_rn(a_dq, dq);
    dt = __dmul_rn(h, pow(__dmul_rn(sf, ratio), 0.3333333));
  }
  const double e100 = 100.0 * 2.220446049250313e-16;
  if (dt < __dadd_rn(__dmul_rn(e100, fabs(t)), e100)) return __dadd_rn(__dmul_rn(e100, t), e100);
  return dt;
}

// predict_q012 + clear_before_step on rows [imin, imax)
__global__ void adv_begin_kernel(double *__restrict__ rec, int imin, int imax, double hnew) {
  const int i = imin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= imax) return;
  adv_begin_row(AdvRowAoS{rec + (size_t)i * AB}, hnew);
}

// predict_position for rows [imin, n) to time t, then pack {q, m} for the grad_v sweeps,
// re-indexed from imin.
__global__ void adv_predict_pack_kernel(double *__restrict__ rec, int imin, int n, double t,
                                        double *__restrict__ packed) {
  const int i = imin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2 xy, zm;
  double cs[3];
  adv_predict_row(AdvRowAoS{rec + (size_t)i * AB}, t, xy, zm, cs);
  double2 *o = reinterpret_cast<double2 *>(packed + (size_t)(i - imin) * PK);
  o[0] = xy;
  o[1] = zm;
  o[2] = make_double2(0.0, 0.0);
  o[3] = make_double2(0.0, 0.0);
}

// dVdqK_x += (vC cs_x[K]) S_x for rows [imin, n)
__global__ void adv_accumulate_kernel(double *__restrict__ rec, int imin, int n, double vC,
                                      const double *__restrict__ S) {
  const int i = imin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const AdvRowAoS b{rec + (size_t)i * AB};
  const double c0 = __dmul_rn(vC, b.ld(32)), c1 = __dmul_rn(vC, b.ld(33)), c2 = __dmul_rn(vC, b.ld(34));
  const double s[3] = {S[3 * (size_t)(i - imin)], S[3 * (size_t)(i - imin) + 1],
                       S[3 * (size_t)(i - imin) + 2]};
  adv_accumulate_row(b, c0, c1, c2, s);
}

// finish_body on rows [imin, imax); on this path h_adaptive stays on the host (libm pow)
__global__ void adv_finish_kernel(double *__restrict__ rec, int imin, int imax, double t) {
  const int i = imin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= imax) return;
  adv_finish_row(AdvRowAoS{rec + (size_t)i * AB}, t);
}

// initialize_before_first_step (advancer.ml:357-393): clear, h := 1, then from the all-pairs
// acc+jerk sums g, dg: dVdq0 = c0 (g - h dg), dVdq1 = c1 (g - h/2 dg), dVdq2 = c2 g.
__global__ void adv_init_kernel(double *__restrict__ rec, int n, const double *__restrict__ g,
                                const double *__restrict__ dg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double *b = rec + (size_t)i * AB;
  const double h = 1.0;
  const double c0 = __ddiv_rn(h, 6.0), c1 = __ddiv_rn(__dmul_rn(2.0, h), 3.0);
  const double c2 = c0;
  b[8] = h;
  b[32] = 1.0;
  b[33] = b[34] = 0.0;
  for (int k = 0; k < 3; ++k) {
    const double gv = g[3 * (size_t)i + k], dgv = dg[3 * (size_t)i + k];
    b[11 + k] = __dmul_rn(c0, __dsub_rn(gv, __dmul_rn(h, dgv)));
    b[14 + k] = __dmul_rn(c1, __dsub_rn(gv, __dmul_rn(__dmul_rn(0.5, h), dgv)));
    b[17 + k] = __dmul_rn(c2, gv);
  }
}

// pack {q, m, p/m} of every record (for the all-pairs sweep of initialize_before_first_step
// and for Energy.energy on the resident state)
__global__ void adv_pack_full_kernel(const double *__restrict__ rec, int n,
                                     double *__restrict__ packed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *b = rec + (size_t)i * AB;
  const double m = b[1];
  double2 *o = reinterpret_cast<double2 *>(packed + (size_t)i * PK);
  o[0] = make_double2(b[2], b[3]);
  o[1] = make_double2(b[4], m);
  o[2] = make_double2(__ddiv_rn(b[5], m), __ddiv_rn(b[6], m));
  o[3] = make_double2(__ddiv_rn(b[7], m), 0.0);
}

// ke[i] = |p_i|^2 / (2 m_i) of every record (energy.ml:57-58), the host formula's sequence
__global__ void adv_kinetic_kernel(const double *__restrict__ rec, int n, double *__restrict__ ke) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *b = rec + (size_t)i * AB;
  const double x = b[5], y = b[6], z = b[7];
  const double n2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
  ke[i] = __ddiv_rn(n2, __dmul_rn(2.0, b[1]));
}

// rows [0, count) := old rows perm[0..count)  (sort_range, advancer.ml:290-297)
__global__ void adv_gather_kernel(const double *__restrict__ src, double *__restrict__ dst,
                                  const int *__restrict__ perm, int count) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)count * AB) return;
  const int row = (int)(e / AB), c = (int)(e % AB);
  dst[e] = src[(size_t)perm[row] * AB + c];
}

// ---- Advancer.A.advance_range ENTIRELY on the device (advancer.ml:316-350) ---------------------
// One persistent cooperative kernel runs the block recursion of `advance_range bs imax dt sf None`
// for a prefix of imax <= ADT_MAX slots: warp 0 of every CTA walks the same explicit frame stack
// (the decisions depend only on the prefix's hmax, which every CTA mirrors in shared memory), and
// each interact_bodies is two grid-synchronous phases:
//   A   every CTA first predicts the ACTIVE block for itself into shared memory (from fields that
//       are stable during the phase: q at a step's first interact, q0..q2/h/t0 afterwards - the
//       same formula on the same inputs as the owner's, hence the same bits).  Then thread-per-
//       slot j in [imin, n): predict_q012 + clear (active slots of a step's first interact),
//       predict_position to t; T_j = sum over the active slots a != j of the pair term; a
//       passive slot is complete and is folded into its dVdq0..2 at once.  The passive slots a
//       CTA owns are also its SOURCE tile (straight from registers to shared memory): thread
//       (a, s) sums tile sources s, s+SPLIT, ... for active a, the SPLIT lanes of a target are
//       combined by a fixed shuffle tree (+ a fixed pass over warps)                | grid barrier
//   B   one warp per active slot adds the CTAs' partial sums in a fixed order and folds T_a + R_a
//       into dVdq0..2.  For a block of a few slots this is left to phase A of the NEXT interact,
//       done by the CTA that owns the slot there (its own barrier then orders the fold before
//       that slot's passive update) - no grid barrier.  A wide block is folded at once by every
//       warp of the grid, followed by a barrier.  After a step's third interact the fold is
//       followed by finish_body and h_adaptive on the same registers.  The T_a / partial-sum
//       buffers alternate between interacts.
// After a step: one more barrier, every CTA reloads the new hmax and sorts its mirror of
// (hmax, slot -> row) with the stable rank of sort_range (advancer.ml:290-297).  The records are
// transposed into a column-major working copy at entry (a warp's loads of one field coalesce)
// and rows never move inside the kernel; at exit they are written back in their sorted order.
// Grids of up to 8 CTAs are launched as one thread-block cluster and use the hardware cluster
// barrier, wider ones are cooperative launches with an atomic counter in L2.  Sums are
// reproducible run to run (fixed grid, fixed trees); their order differs from the tiled sweeps',
// so this path agrees with the per-phase path to rounding (1e-13), not bit for bit.
constexpr int ADT_MAX = 1024;    // largest prefix (and active block) the device recursion takes
constexpr int ADT_DEPTH = 192;   // frames of the explicit recursion stack
constexpr int ADT_TS = 512;      // source tile = one slot per thread of the CTA

struct AdvTreeParams {
  double *rec;                 // [n][AB]  the records (ABI layout): read at entry, rewritten at exit
  double *soa;                 // [AB][ld] column-major working copy, rows where they started
  size_t ld;
  double *hmax;                // [imax]  in: the prefix's hmax (then the steps' new values in transit)
  double *hmax_out;            // [imax]  out: hmax in the final slot order
  int *map;                    // [imax]  out: slot -> row it started in
  double *tact;                // [2][ADT_MAX][6]  T_a and the weights vC cs of the active slots
  double *partial;             // [2][gridDim.x][ADT_MAX][3]   (buffers alternate between interacts)
  unsigned long long *bar;     // grid barrier counter (zeroed before the launch)
  long long *out;              // [16] interact_bodies calls, pair interactions, error, block steps,
                               // then CTA 0's clock cycles per phase: control, P1, barrier, P2 A,
                               // P2 B, barrier, P3, barrier, sort
  int n, imax;
  int cluster;  // the grid is one thread-block cluster (<= 8 CTAs)
  double dt, sf, eps2;
};
struct AdvFrame {
  int imax, imin, stage, pad;
  double dt, t0;
};
struct AdvCtl {
  int op, imin, imax, begin, last, pad;
  double t, vC, hnew;
};
constexpr size_t ADT_SMEM = sizeof(double) * (ADT_MAX * 4 + ADT_TS * 4 + 2 * ADT_MAX + 64 + ADT_MAX * 3) +
                            sizeof(AdvFrame) * ADT_DEPTH + sizeof(int) * 2 * ADT_MAX;

__device__ __forceinline__ void adv_cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ double adv_sort_key(double h) { return h != h ? __longlong_as_double(0x7ff0000000000000LL) : h; }

// one pair: target at (x,y,z), source {sx,sy,sz,sm}; acc += sm d / s^3, d = source - target
__device__ __forceinline__ void adv_pair(double x, double y, double z, double2 sxy, double2 szm,
                                         double eps2, bool self, double &ax, double &ay, double &az) {
  const double dx = sxy.x - x, dy = sxy.y - y, dz = szm.x - z;
  double r2 = fma(dx, dx, eps2);
  r2 = fma(dy, dy, r2);
  r2 = fma(dz, dz, r2);
  double mj = szm.y;
  r2 = self ? 1.0 : r2;
  mj = self ? 0.0 : mj;
  const double yv = rsqrt_nr(r2);
  const double w = (mj * yv) * (yv * yv);
  ax = fma(w, dx, ax);
  ay = fma(w, dy, ay);
  az = fma(w, dz, az);
}

// Folds an active slot's sums into its dVdq0..2: T_a (+ weights vC cs) from `ta`, the CTAs'
// partial sums R_a from part[c * ADT_MAX * 3 + 0..2], c < gc, added in a fixed order by the warp.
// Lane 0 owns the row: everything it needs is requested before the partial sums, so the fold is
// one trip to L2.  last: finish_body and h_adaptive follow on the values (advancer.ml:339-342).
template <class Row>
__device__ __forceinline__ void adv_fold_row(const Row b, const double *ta_g, const double *part,
                                             int gc, int lane, bool last, double t, double sf,
                                             double *hmax_out) {
  double ta[6], v[9], p0[3], q0[3], q2[3], m = 0.0, h = 0.0;
  if (lane == 0) {
    for (int k = 0; k < 6; ++k) ta[k] = __ldcg(ta_g + k);
    for (int k = 0; k < 9; ++k) v[k] = b.ld(11 + k);
    m = b.ld(1);
    if (last) {
      h = b.ld(8);
      for (int k = 0; k < 3; ++k) {
        p0[k] = b.ld(20 + k);
        q0[k] = b.ld(23 + k);
        q2[k] = b.ld(29 + k);
      }
    }
  }
  double r0 = 0.0, r1 = 0.0, r2 = 0.0;
  for (int c = lane; c < gc; c += 32) {
    const double *pi = part + (size_t)c * ADT_MAX * 3;
    r0 += __ldcg(pi);
    r1 += __ldcg(pi + 1);
    r2 += __ldcg(pi + 2);
  }
  for (int off = 16; off > 0; off >>= 1) {
    r0 += __shfl_xor_sync(0xffffffffu, r0, off);
    r1 += __shfl_xor_sync(0xffffffffu, r1, off);
    r2 += __shfl_xor_sync(0xffffffffu, r2, off);
  }
  if (lane == 0) {
    const double sm = -m;
    const double s[3] = {sm * (ta[0] + r0), sm * (ta[1] + r1), sm * (ta[2] + r2)};
    for (int k = 0; k < 3; ++k) {  // adv_accumulate_row on the values
      v[k] = __dadd_rn(v[k], __dmul_rn(ta[3], s[k]));
      v[3 + k] = __dadd_rn(v[3 + k], __dmul_rn(ta[4], s[k]));
      v[6 + k] = __dadd_rn(v[6 + k], __dmul_rn(ta[5], s[k]));
    }
    for (int k = 0; k < 9; ++k) b.st(11 + k, v[k]);
    if (last) {
      double q[3], p[3];
      adv_finish_vals(h, m, v, p0, q0, q, p);
      for (int k = 0; k < 3; ++k) {
        b.st(2 + k, q[k]);
        b.st(5 + k, p[k]);
      }
      b.st(0, t);
      *hmax_out = adv_h_adaptive_vals(h, m, t, q, q2, v + 6, sf);
    }
  }
}

template <int NTMAX>  // register budget: 255 per thread up to 256 threads, 128 at 512
__global__ void __launch_bounds__(NTMAX, 1) adv_tree_kernel(const AdvTreeParams P) {
  extern __shared__ __align__(16) unsigned char adt_smem[];
  double *xa = reinterpret_cast<double *>(adt_smem);  // active {q, m}
  double *tile = xa + ADT_MAX * 4;                    // passive source tile
  double *hs = tile + ADT_TS * 4;                     // hmax of the prefix, slot order
  double *hs2 = hs + ADT_MAX;
  double *red = hs2 + ADT_MAX;                        // [16 warps][3]
  double *racc = red + 64;                            // [ADT_MAX][3] this CTA's sums for the active slots
  AdvFrame *stack = reinterpret_cast<AdvFrame *>(racc + ADT_MAX * 3);
  int *ms = reinterpret_cast<int *>(stack + ADT_DEPTH);  // slot -> row
  int *ms2 = ms + ADT_MAX;
  __shared__ AdvCtl ctl;
  __shared__ int s_sp;

  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, G = gridDim.x;
  const int gt = blockIdx.x * T + tid, GT = G * T;
  const int n = P.n, top = P.imax;
  const double eps2 = P.eps2;
  unsigned long long nbar = 0;
  // up to 8 CTAs are launched as ONE thread-block cluster: the hardware cluster barrier (release /
  // acquire at cluster scope) replaces the atomic counter in L2
#define ADT_SYNC()                                                       \
  do {                                                                   \
    if (P.cluster) adv_cluster_barrier();                                \
    else adv_grid_barrier(P.bar, (unsigned long long)G * ++nbar);        \
  } while (0)
  int n_int = 0;  // interacts so far (its parity picks the T_a / partial-sum buffers)
  __shared__ long long s_cnt[2];  // thread 0's counters: pair interactions, block steps
  int pend_imin = 0, pend_na = 0, pend_gc = 0;  // the previous interact's fold, still to be done
  const double *pend_tact = nullptr, *pend_part = nullptr;
  __shared__ long long s_cyc[10];  // thread 0's cycle counters per phase, [9] = the last reading
  if (tid == 0) {
    for (int k = 0; k < 9; ++k) s_cyc[k] = 0;
    s_cyc[9] = clock64();
    s_cnt[0] = s_cnt[1] = 0;
  }
#define ADT_TICK(k)                       \
  do {                                    \
    if (tid == 0) {                       \
      const long long c_now = clock64();  \
      s_cyc[k] += c_now - s_cyc[9];       \
      s_cyc[9] = c_now;                   \
    }                                     \
  } while (0)

  for (int i = tid; i < top; i += T) {
    hs[i] = __ldcg(P.hmax + i);
    ms[i] = i;
  }
  if (tid == 0) {
    stack[0] = AdvFrame{top, 0, 0, 0, P.dt, 0.0};
    s_sp = 1;
  }
  const size_t LD = P.ld;
  for (size_t e = gt; e < (size_t)n * AB; e += GT) {  // records -> column-major working copy
    const size_t row = e / AB;
    P.soa[(e - row * AB) * LD + row] = __ldcg(P.rec + e);
  }
  ADT_SYNC();

  for (;;) {
    // ---- control: warp 0 walks the frames until the next interact_bodies (lane 0 writes) -----
    if (tid < 32) {
      int sp = s_sp, op = 0;
      while (sp > 0) {
        __syncwarp();
        const AdvFrame f = stack[sp - 1];
        __syncwarp();
        if (f.stage == 0) {
          if (hs[f.imax - 1] < f.dt) {  // advancer.ml:317-319: two half steps
            if (sp >= ADT_DEPTH) { op = -1; break; }
            if (lane == 0) {
              stack[sp - 1].stage = 10;
              stack[sp] = AdvFrame{f.imax, 0, 0, 0, f.dt / 2.0, 0.0};
            }
            ++sp;
            continue;
          }
          int imin = 0;  // advancer.ml:321-328, scanned 32 slots at a time from the top
          for (int hi = f.imax; hi > 0; hi -= 32) {
            const int i = hi - 1 - lane;
            const bool stop = i >= 0 && !(f.dt < hs[i]);
            const unsigned bal = __ballot_sync(0xffffffffu, stop);
            if (bal) {
              imin = hi - (__ffs(bal) - 1);
              break;
            }
          }
          if (imin >= n) { op = -2; break; }  // bs.(imin).t out of bounds (the reference raises)
          const double t0 = __ldcg(P.soa + (imin < top ? ms[imin] : imin));  // let t0 = bs.(imin).t
          if (imin > 0 && sp >= ADT_DEPTH) { op = -1; break; }
          if (lane == 0) {
            stack[sp - 1].imin = imin;
            stack[sp - 1].t0 = t0;
            stack[sp - 1].stage = 1;
            ctl = AdvCtl{1, imin, f.imax, 1, 0, 0, t0, f.dt / 6.0, f.dt};
            if (imin > 0) stack[sp] = AdvFrame{imin, 0, 0, 0, f.dt / 2.0, 0.0};
          }
          if (imin > 0) ++sp;
          op = 1;
          break;
        } else if (f.stage == 1) {
          if (f.imin > 0 && sp >= ADT_DEPTH) { op = -1; break; }
          if (lane == 0) {
            stack[sp - 1].stage = 2;
            ctl = AdvCtl{1, f.imin, f.imax, 0, 0, 0, f.t0 + 0.5 * f.dt, 2.0 * f.dt / 3.0, f.dt};
            if (f.imin > 0) stack[sp] = AdvFrame{f.imin, 0, 0, 0, f.dt / 2.0, 0.0};
          }
          if (f.imin > 0) ++sp;
          op = 1;
          break;
        } else if (f.stage == 2) {
          if (lane == 0) ctl = AdvCtl{1, f.imin, f.imax, 0, 1, 0, f.t0 + f.dt, f.dt / 6.0, f.dt};
          --sp;
          op = 1;
          break;
        } else if (f.stage == 10) {  // second half step
          if (sp >= ADT_DEPTH) { op = -1; break; }
          if (lane == 0) {
            stack[sp - 1].stage = 11;
            stack[sp] = AdvFrame{f.imax, 0, 0, 0, f.dt / 2.0, 0.0};
          }
          ++sp;
        } else {
          --sp;
        }
      }
      if (lane == 0) {
        s_sp = sp;
        if (op != 1) ctl.op = op;
      }
    }
    __syncthreads();
    ADT_TICK(0);
    if (ctl.op != 1) break;
    const int imin = ctl.imin, imax = ctl.imax, na = imax - imin;
    const bool begin = ctl.begin != 0, last = ctl.last != 0;
    const double t = ctl.t, vC = ctl.vC, hnew = ctl.hnew;
    n_int += 1;
    if (tid == 0) {
      s_cnt[0] += (long long)na * (n - imin - 1) - (long long)na * (na - 1) / 2;
      s_cnt[1] += last ? 1 : 0;
    }

    double *tactb = P.tact + (size_t)(n_int & 1) * ADT_MAX * 6;
    double *partb = P.partial + (size_t)(n_int & 1) * G * ADT_MAX * 3;
    const int Gc = min(G, (n - imin + T - 1) / T);  // CTAs that own a slot

    // ---- A.0: the previous interact's fold, by the CTA that owns the slot in THIS phase (so it
    // is ordered before that slot's passive update below by the CTA's own barrier) -------------
    if (pend_na > 0) {
      for (int a = tid >> 5; a < pend_na; a += T >> 5) {
        const int slot = pend_imin + a;
        if (((slot - imin) / T) % G != (int)blockIdx.x) continue;
        adv_fold_row(AdvRowSoA{P.soa + ms[slot], LD}, pend_tact + (size_t)a * 6,
                     pend_part + (size_t)a * 3, pend_gc, lane, false, 0.0, 0.0, nullptr);
      }
      pend_na = 0;
    }

    // ---- A.1: the active block, predicted by every CTA for itself ----------------------------
    for (int a = tid; a < na; a += T) {
      const AdvRowSoA b{P.soa + ms[imin + a], LD};
      double x[3];
      const double m = b.ld(1);
      if (begin) {  // bs.(a).t = t0 here: predict_position leaves q alone (advancer.ml:174-175)
        for (int k = 0; k < 3; ++k) x[k] = b.ld(2 + k);
      } else {
        const double t0 = b.ld(10), h = b.ld(8);
        double q0[3], q1[3], q2[3], c[3];
        for (int k = 0; k < 3; ++k) {
          q0[k] = b.ld(23 + k);
          q1[k] = b.ld(26 + k);
          q2[k] = b.ld(29 + k);
        }
        adv_predict_vals(t, t0, h, q0, q1, q2, c, x);
      }
      xa[4 * a] = x[0];
      xa[4 * a + 1] = x[1];
      xa[4 * a + 2] = x[2];
      xa[4 * a + 3] = m;
      racc[3 * a] = racc[3 * a + 1] = racc[3 * a + 2] = 0.0;
    }
    ADT_TICK(1);

    // ---- A.2: passes of one slot per thread ---------------------------------------------------
    for (int r = 0;; ++r) {
      const long long j0l = (long long)imin + (long long)blockIdx.x * T + (long long)r * GT;
      if (j0l >= n) break;
      const int j0 = (int)j0l, j = j0 + tid;
      const bool valid = j < n;
      const AdvRowSoA b{P.soa + (valid ? (j < top ? ms[j] : j) : 0), LD};
      double2 mxy = make_double2(0.0, 0.0), mzm = mxy;
      double c0 = 0.0, c1 = 0.0, c2 = 0.0, v[9];
      __syncthreads();  // xa / racc are set; the previous pass is done with the tile
      if (valid) {
        if (j >= imax)
          for (int k = 0; k < 9; ++k) v[k] = b.ld(11 + k);
        if (begin && j < imax) adv_begin_row(b, hnew);
        double cs[3];
        adv_predict_row(b, t, mxy, mzm, cs);
        c0 = __dmul_rn(vC, cs[0]);
        c1 = __dmul_rn(vC, cs[1]);
        c2 = __dmul_rn(vC, cs[2]);
        reinterpret_cast<double2 *>(tile)[2 * tid] = mxy;
        reinterpret_cast<double2 *>(tile)[2 * tid + 1] = mzm;
      }
      __syncthreads();
      ADT_TICK(2);
      if (valid && na > 0) {  // this slot against the active block
        double ax = 0.0, ay = 0.0, az = 0.0;
        const int aself = j - imin;  // >= na for a passive slot: never matches
        const double2 *x2 = reinterpret_cast<const double2 *>(xa);
#pragma unroll 4
        for (int a = 0; a < na; ++a)
          adv_pair(mxy.x, mxy.y, mzm.x, x2[2 * a], x2[2 * a + 1], eps2, a == aself, ax, ay, az);
        if (j >= imax) {
          const double sm = -mzm.y;
          const double s[3] = {sm * ax, sm * ay, sm * az};
          adv_accumulate_row_pre(b, v, c0, c1, c2, s);
        } else {
          double *o = tactb + (size_t)aself * 6;
          o[0] = ax;
          o[1] = ay;
          o[2] = az;
          o[3] = c0;
          o[4] = c1;
          o[5] = c2;
        }
      }
      ADT_TICK(3);
      // the active block against the passive slots of this pass (tile entries [klo, khi))
      const int klo = max(0, imax - j0), khi = min(T, n - j0);
      if (klo < khi) {
        for (int a0 = 0; a0 < na; a0 += T) {
          const int nac = min(T, na - a0);
          int napad = 1;
          while (napad < nac) napad <<= 1;
          const int SPLIT = T / napad;
          const int al = tid / SPLIT, s = tid - al * SPLIT;
          const bool live = al < nac;
          double r0 = 0.0, r1 = 0.0, r2 = 0.0;
          if (live) {
            const double x = xa[4 * (a0 + al)], y = xa[4 * (a0 + al) + 1], z = xa[4 * (a0 + al) + 2];
            const double2 *t2 = reinterpret_cast<const double2 *>(tile);
            for (int k = klo + s; k < khi; k += SPLIT)
              adv_pair(x, y, z, t2[2 * k], t2[2 * k + 1], eps2, false, r0, r1, r2);
          }
          const int width = SPLIT < 32 ? SPLIT : 32;
          for (int off = width >> 1; off > 0; off >>= 1) {
            r0 += __shfl_xor_sync(0xffffffffu, r0, off);
            r1 += __shfl_xor_sync(0xffffffffu, r1, off);
            r2 += __shfl_xor_sync(0xffffffffu, r2, off);
          }
          if (SPLIT > 32) {  // several warps per target: fixed pass over them
            __syncthreads();
            if (lane == 0) {
              red[3 * (tid >> 5)] = r0;
              red[3 * (tid >> 5) + 1] = r1;
              red[3 * (tid >> 5) + 2] = r2;
            }
            __syncthreads();
            if (tid < nac) {
              const int wpt = SPLIT >> 5;
              double q0 = 0.0, q1 = 0.0, q2 = 0.0;
              for (int w = 0; w < wpt; ++w) {
                q0 += red[3 * (tid * wpt + w)];
                q1 += red[3 * (tid * wpt + w) + 1];
                q2 += red[3 * (tid * wpt + w) + 2];
              }
              racc[3 * (a0 + tid)] += q0;
              racc[3 * (a0 + tid) + 1] += q1;
              racc[3 * (a0 + tid) + 2] += q2;
            }
          } else if (live && s == 0) {
            racc[3 * (a0 + al)] += r0;
            racc[3 * (a0 + al) + 1] += r1;
            racc[3 * (a0 + al) + 2] += r2;
          }
        }
      }
    }
    __syncthreads();
    if ((int)blockIdx.x < Gc)
      for (int e = tid; e < 3 * na; e += T) partb[(size_t)blockIdx.x * ADT_MAX * 3 + e] = racc[e];
    ADT_TICK(4);
    ADT_SYNC();
    ADT_TICK(5);
    if (na == 0) continue;  // hmax = dt exactly: an empty block (advancer.ml:321-328), nothing to fold

    // ---- B: after a step's third interact, one warp per active slot: fold, finish_body,
    // h_adaptive (advancer.ml:339-342).  Other interacts leave the fold to the next phase A.
    if (!last && na <= (T >> 4)) {  // a few slots: their owners' CTAs fold them in the next phase A
      pend_imin = imin;
      pend_na = na;
      pend_gc = Gc;
      pend_tact = tactb;
      pend_part = partb;
      continue;
    }
    if (!last) {  // a wide block: every warp of the grid takes slots, then a barrier orders the
                  // folds before the next interact's passive updates of the same rows
      for (int a = gt >> 5; a < na; a += GT >> 5)
        adv_fold_row(AdvRowSoA{P.soa + ms[imin + a], LD}, tactb + (size_t)a * 6,
                     partb + (size_t)a * 3, Gc, lane, false, 0.0, 0.0, nullptr);
      ADT_TICK(6);
      ADT_SYNC();
      ADT_TICK(7);
      continue;
    }
    for (int a = gt >> 5; a < na; a += GT >> 5)
      adv_fold_row(AdvRowSoA{P.soa + ms[imin + a], LD}, tactb + (size_t)a * 6,
                   partb + (size_t)a * 3, Gc, lane, true, t, P.sf, P.hmax + imin + a);
    pend_na = 0;
    ADT_TICK(6);
    if (last) {
      ADT_SYNC();
      ADT_TICK(7);
      for (int i = imin + tid; i < imax; i += T) hs[i] = __ldcg(P.hmax + i);
      __syncthreads();
      // sort_range bs hmax_lt 0 imax (advancer.ml:290-297, 349): stable rank, NaN keyed as +inf.
      // Only the block [imin, imax) has new keys; when the rest of the prefix is still in order
      // (it is, unless NaNs got in) a slot's rank needs the block's keys and one binary search.
      int bad = 0, badlow = 0;
      for (int i = tid; i + 1 < imax; i += T) {
        const int b1 = adv_sort_key(hs[i + 1]) < adv_sort_key(hs[i]);
        bad |= b1;
        badlow |= (i + 1 < imin) ? b1 : 0;
      }
      const int flags = __syncthreads_or(bad | (badlow << 1));
      if (flags & 1) {
        for (int i = tid; i < imax; i += T) {
          const double ki = adv_sort_key(hs[i]);
          int r = 0;
          if (!(flags & 2)) {
            if (i < imin) {  // ties: a slot below the block stays ahead of the block's slots
              r = i;
              for (int k = imin; k < imax; ++k) r += adv_sort_key(hs[k]) < ki ? 1 : 0;
            } else {
              int lo = 0, hi = imin;  // slots below the block with key <= ki
              while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (adv_sort_key(hs[mid]) <= ki) lo = mid + 1;
                else hi = mid;
              }
              r = lo;
              for (int k = imin; k < imax; ++k) {
                const double kk = adv_sort_key(hs[k]);
                r += (kk < ki || (kk == ki && k < i)) ? 1 : 0;
              }
            }
          } else {
            for (int k = 0; k < imax; ++k) {
              const double kk = adv_sort_key(hs[k]);
              r += (kk < ki || (kk == ki && k < i)) ? 1 : 0;
            }
          }
          hs2[r] = hs[i];
          ms2[r] = ms[i];
        }
        __syncthreads();
        for (int i = tid; i < imax; i += T) {
          hs[i] = hs2[i];
          ms[i] = ms2[i];
        }
        __syncthreads();
      }
      ADT_TICK(8);
    }
  }

  // working copy -> records, each row to the slot the sorts gave it (every write above is behind
  // a grid barrier by now)
  for (size_t e = gt; e < (size_t)n * AB; e += GT) {
    const size_t row = e / AB;
    const size_t src = row < (size_t)top ? (size_t)ms[row] : row;
    P.rec[e] = __ldcg(P.soa + (e - row * AB) * LD + src);
  }
  if (blockIdx.x == 0) {
    for (int i = tid; i < top; i += T) {
      P.hmax_out[i] = hs[i];
      P.map[i] = ms[i];
    }
    if (tid == 0) {
      P.out[0] = n_int;
      P.out[1] = s_cnt[0];
      P.out[2] = ctl.op;  // 0 done, -1 recursion deeper than ADT_DEPTH, -2 empty block at the array's end
      P.out[3] = s_cnt[1];
      for (int k = 0; k < 9; ++k) P.out[4 + k] = s_cyc[k];
    }
  }
}

// ---- device-resident Leapfrog.b array (leapfrog.ml:5-12): packed rows + dt_max ---------------
// sort_sub (leapfrog.ml:106-111): rows [0,count) := old rows perm[0..count)
__global__ void lf_gather_kernel(const double *__restrict__ src, double *__restrict__ dst,
                                 const double *__restrict__ dsrc, double *__restrict__ ddst,
                                 const int *__restrict__ perm, int count) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)count * PK) return;
  const int row = (int)(e / PK), c = (int)(e % PK);
  dst[e] = src[(size_t)perm[row] * PK + c];
  if (c == 0) ddst[row] = dsrc[perm[row]];
}
// sort_sub (leapfrog.ml:106-111) on the device: Array.fast_sort (stable) of the prefix by
// `compare b1.dt_max b2.dt_max` - OCaml's polymorphic compare on floats: nan equals nan and sorts
// below everything else, -0.0 equals 0.0.  Keys are first mapped to unsigned integers in that
// order (lf_sort_key_kernel); then rank by counting: new position of row i =
// #{j < i : key_j <= key_i} + #{j > i : key_j < key_i} - iend^2 integer comparisons spread over a
// 2-D grid (256 rows x a chunk of the keys per CTA, chunk staged in shared memory, partial counts
// added atomically: integers, so the result does not depend on the order); lf_scatter_perm_kernel
// then writes perm[rank] = i (new row s = old row perm[s]).
__global__ void lf_sort_key_kernel(const double *__restrict__ key, int iend,
                                   unsigned long long *__restrict__ ukey, int *__restrict__ rank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= iend) return;
  const double k = key[i];
  unsigned long long u;
  if (k != k) u = 0ULL;                                    // nan: below everything, all equal
  else if (k == 0.0) u = 0x8000000000000000ULL;            // -0.0 = 0.0
  else {
    const unsigned long long b = (unsigned long long)__double_as_longlong(k);
    u = (b >> 63) ? ~b : (b | 0x8000000000000000ULL);     // negatives reversed, positives above
    if (u == 0ULL) u = 1ULL;                               // (unreachable: ~b = 0 needs b = all ones = nan)
  }
  ukey[i] = u;
  rank[i] = 0;
}
constexpr int LFS_CHUNK = 2048;  // keys per CTA in the j direction (16 KB of shared memory)
__global__ void __launch_bounds__(256) lf_rank_kernel(const unsigned long long *__restrict__ ukey,
                                                      int iend, int *__restrict__ rank) {
  __shared__ unsigned long long tile[LFS_CHUNK];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int j0 = blockIdx.y * LFS_CHUNK, cnt = min(LFS_CHUNK, iend - j0);
  for (int t = threadIdx.x; t < cnt; t += 256) tile[t] = ukey[j0 + t];
  __syncthreads();
  if (i >= iend) return;
  const unsigned long long ki = ukey[i];
  int r = 0;
  // j < i: ties count (stable); j > i: strictly smaller only; j == i contributes nothing
  const int before = min(max(i - j0, 0), cnt);       // tile entries with j < i
  const int after0 = min(max(i + 1 - j0, 0), cnt);   // first tile entry with j > i
#pragma unroll 4
  for (int t = 0; t < before; ++t) r += tile[t] <= ki ? 1 : 0;
#pragma unroll 4
  for (int t = after0; t < cnt; ++t) r += tile[t] < ki ? 1 : 0;
  if (r) atomicAdd(rank + i, r);
}
__global__ void lf_scatter_perm_kernel(const int *__restrict__ rank, int iend, int *__restrict__ perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < iend) perm[rank[i]] = i;
}
// drift dt (leapfrog.ml:64-68) on rows [i0, i1)
__global__ void lf_drift_kernel(double *__restrict__ packed, int i0, int i1, double dt) {
  int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= i1) return;
  double *b = packed + (size_t)i * PK;
  b[0] = __dadd_rn(b[0], __dmul_rn(dt, b[4]));
  b[1] = __dadd_rn(b[1], __dmul_rn(dt, b[5]));
  b[2] = __dadd_rn(b[2], __dmul_rn(dt, b[6]));
}
// after a kick block (leapfrog.ml:125-137 for bodies [0, iend), active block [ibegin, iend)):
// v += dv; dt_max = min(dt_max, tmin) as leapfrog.ml:103-104 writes it, where an active body
// starts from infinity (the reset of :125-127, which nothing reads before this point); then the
// second half drift of the active bodies.
__global__ void lf_apply_kernel(double *__restrict__ packed, double *__restrict__ dt_max,
                                const double *__restrict__ dv, const double *__restrict__ tmin,
                                int ibegin, int iend, int with_drift, double drift_after) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= iend) return;
  double *b = packed + (size_t)i * PK;
  b[4] = __dadd_rn(b[4], dv[3 * i]);
  b[5] = __dadd_rn(b[5], dv[3 * i + 1]);
  b[6] = __dadd_rn(b[6], dv[3 * i + 2]);
  if (tmin != nullptr) {
    double d = i >= ibegin ? __longlong_as_double(0x7ff0000000000000LL) : dt_max[i];
    if (d > tmin[i]) d = tmin[i];
    dt_max[i] = d;
  } else if (i >= ibegin) {
    dt_max[i] = __longlong_as_double(0x7ff0000000000000LL);
  }
  if (with_drift && i >= ibegin) {
    b[0] = __dadd_rn(b[0], __dmul_rn(drift_after, b[4]));
    b[1] = __dadd_rn(b[1], __dmul_rn(drift_after, b[5]));
    b[2] = __dadd_rn(b[2], __dmul_rn(drift_after, b[6]));
  }
}

// ---- Leapfrog.advance_subsystem ENTIRELY on the device (leapfrog.ml:113-142) -------------------
// `advance_subsystem dt eta bs iend` touches nothing but bodies [0, iend): a subtree of the
// block recursion with iend <= LFT_MAX is self-contained, and one CTA runs all of it with the
// bodies in shared memory - explicit frame stack, find_can_advance_index, drifts, the kick block
// of every step (same pair function and FP32 filter as the tiled kick sweep: velocity kicks from
// the pre-kick state, bit-exact pair time-scales), dt_max updates, stable sort_sub.  Rows stay
// where they are; the sorts permute a slot -> row map (and dt_max, t), applied when the rows are
// written back.  Kick block, targets [0, iend) with active block [ibegin, iend):
// an active slot sums every other slot of [0, iend), a passive one the active block; SPLIT lanes
// share a target and take sources s, s+SPLIT, ..., combined by a fixed shuffle tree (sum of dv,
// min of the time-scale).
constexpr int LFT_MAX = 1024;
constexpr int LFT_THREADS = 512;
constexpr int LFT_DEPTH = 160;
constexpr int LFT_CLUSTER = 8;  // CTAs of the cluster variant (wide subtrees)

struct LfTreeParams {
  double *packed;   // [iend][PK] in/out, array order
  double *dt_max;   // [iend] in/out
  double *t;        // [iend] in/out (the bodies' times, drifted like leapfrog.ml:65)
  int *map;         // [iend] out: new row s = old row map[s]
  long long *out;   // kick blocks, error
  int iend;
  double dt, eta;
};
struct LfFrame {
  int iend, ibegin, stage, pad;
  double dt;
};
constexpr size_t LFT_SMEM = sizeof(double) * ((size_t)LFT_MAX * PK + 4 * LFT_MAX + 4 * LFT_MAX + 64) +
                            sizeof(LfFrame) * LFT_DEPTH + sizeof(int) * 2 * LFT_MAX;

// OCaml's compare on floats puts nan below everything (leapfrog.ml:108)
__device__ __forceinline__ double lf_sort_key(double d) {
  return d != d ? __longlong_as_double(0xfff0000000000000LL) : d;
}

// NC = 1: one CTA.  NC = LFT_CLUSTER: the grid is ONE thread-block cluster; every CTA keeps a full
// replica of the state and runs the same control flow, but a kick block's targets are dealt out
// (slot x belongs to CTA x mod NC): a CTA sums, applies and drifts its own slots, then - behind a
// cluster barrier, once nobody reads the old rows any more - stores the updated rows, dt_max and
// t into the other replicas through distributed shared memory, and a second cluster barrier
// publishes them.  Wide blocks run NC times faster; the sums per slot are the same in both.
template <bool TS, int NC>
__global__ void __launch_bounds__(LFT_THREADS, 1) lf_tree_kernel(const LfTreeParams P) {
  using Op = OpKick<TS>;
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = NC > 1 ? (int)cluster.block_rank() : 0;
  extern __shared__ __align__(16) unsigned char lft_smem[];
  double *rows = reinterpret_cast<double *>(lft_smem);  // [LFT_MAX][PK]
  double *dtm = rows + (size_t)LFT_MAX * PK;           // dt_max, slot order
  double *tt = dtm + LFT_MAX;                          // t, slot order
  double *dtm2 = tt + LFT_MAX;
  double *tt2 = dtm2 + LFT_MAX;
  double *res = tt2 + LFT_MAX;                         // [LFT_MAX][4] dv (unscaled), tmin per slot
  double *red = res + 4 * LFT_MAX;                     // [16 warps][4]
  LfFrame *stack = reinterpret_cast<LfFrame *>(red + 64);
  int *ms = reinterpret_cast<int *>(stack + LFT_DEPTH);  // slot -> row
  int *ms2 = ms + LFT_MAX;

  const int T = LFT_THREADS, tid = threadIdx.x, lane = tid & 31;
  const int top = P.iend;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  SweepParams SP;
  SP.eta = P.eta;
  long long nk = 0;
  int sp = 0, err = 0;

  for (int e = tid; e < top * PK; e += T) rows[e] = P.packed[e];
  for (int i = tid; i < top; i += T) {
    dtm[i] = P.dt_max[i];
    tt[i] = P.t[i];
    ms[i] = i;
  }
  if (tid == 0) stack[0] = LfFrame{top, 0, 0, 0, P.dt};
  sp = 1;
  __syncthreads();

  // every thread walks the same frames (the data it looks at is shared and only changes between
  // barriers); thread 0 alone writes the stack
  while (sp > 0 && !err) {
    const LfFrame f = stack[sp - 1];
    __syncthreads();
    if (f.stage == 0) {
      if (f.iend <= 0) {  // leapfrog.ml:120-121
        --sp;
        continue;
      }
      int ibegin = f.iend;  // find_can_advance_index (leapfrog.ml:113-117)
      while (!(ibegin == 0 || dtm[ibegin - 1] < f.dt)) --ibegin;
      __syncthreads();
      for (int i = ibegin + tid; i < f.iend; i += T) dtm[i] = INF;  // leapfrog.ml:125-127
      if (sp >= LFT_DEPTH) {
        err = 1;
        break;
      }
      if (tid == 0) {
        stack[sp - 1].ibegin = ibegin;
        stack[sp - 1].stage = 1;
        stack[sp] = LfFrame{ibegin, 0, 0, 0, f.dt / 2.0};
      }
      ++sp;
      __syncthreads();
      continue;
    }
    if (f.stage == 2) {  // sort_sub bs iend (leapfrog.ml:106-111, 141): stable rank
      const int iend = f.iend;
      int bad = 0;
      for (int i = tid; i + 1 < iend; i += T) bad |= lf_sort_key(dtm[i + 1]) < lf_sort_key(dtm[i]);
      if (__syncthreads_or(bad)) {
        for (int i = tid; i < iend; i += T) {
          const double ki = lf_sort_key(dtm[i]);
          int r = 0;
          for (int k = 0; k < iend; ++k) {
            const double kk = lf_sort_key(dtm[k]);
            r += (kk < ki || (kk == ki && k < i)) ? 1 : 0;
          }
          dtm2[r] = dtm[i];
          tt2[r] = tt[i];
          ms2[r] = ms[i];
        }
        __syncthreads();
        for (int i = tid; i < iend; i += T) {
          dtm[i] = dtm2[i];
          tt[i] = tt2[i];
          ms[i] = ms2[i];
        }
      }
      --sp;
      __syncthreads();
      continue;
    }
    // stage 1: drift, kick block, drift (leapfrog.ml:129-139), then the second half recursion
    const int ibegin = f.ibegin, iend = f.iend, na = iend - ibegin;
    const double dt = f.dt, dt2 = f.dt / 2.0;
    if (na > 0) {
      for (int i = ibegin + tid; i < iend; i += T) {
        double *b = rows + (size_t)ms[i] * PK;
        tt[i] = __dadd_rn(tt[i], dt2);
        b[0] = __dadd_rn(b[0], __dmul_rn(dt2, b[4]));
        b[1] = __dadd_rn(b[1], __dmul_rn(dt2, b[5]));
        b[2] = __dadd_rn(b[2], __dmul_rn(dt2, b[6]));
      }
      __syncthreads();
      // the kick sums of this CTA's slots x = rank, rank + NC, ...: an active slot against every
      // other slot of [0, iend), a passive one against the active block.  SPLIT lanes share a
      // target (sources lo + s, lo + s + SPLIT, ...), combined by a fixed shuffle tree (sum of
      // dv, min of the time-scale) and, beyond a warp, a fixed pass over the warps.  The running
      // minimum starts from the slot's current dt_max - nothing above it matters (:103-104) - so
      // the FP32 filter rejects most pairs of a passive slot at once.
      const int nslots = rank < iend ? (iend - rank + NC - 1) / NC : 0;
      for (int u0 = 0; u0 < nslots; u0 += T) {
        const int nac = min(T, nslots - u0);
        int napad = 1;
        while (napad < nac) napad <<= 1;
        const int SPLIT = T / napad;
        const int ul = tid / SPLIT, sl = tid - ul * SPLIT;
        const bool live = ul < nac;
        const int x = rank + NC * (u0 + ul);
        double acc[Op::NACC < 5 ? 5 : Op::NACC];
        Op::zero(acc);
        if (live) {
          typename Op::Tgt tg;
          Op::load(tg, rows + (size_t)ms[x] * PK, SP, 0);
          if (TS) {
            acc[3] = dtm[x];  // infinity for an active slot (reset in stage 0)
            acc[4] = Op::filter_state(acc[3], SP.eta);
            acc[5] = acc[3];
          }
          const int lo = x >= ibegin ? 0 : ibegin;
          for (int k = lo + sl; k < iend; k += SPLIT)
            Op::template pair<true>(tg, acc, reinterpret_cast<const double2 *>(rows + (size_t)ms[k] * PK),
                                    nullptr, k == x, SP);
        }
        const int width = SPLIT < 32 ? SPLIT : 32;
        for (int off = width >> 1; off > 0; off >>= 1) {
          acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], off);
          acc[1] += __shfl_xor_sync(0xffffffffu, acc[1], off);
          acc[2] += __shfl_xor_sync(0xffffffffu, acc[2], off);
          if (TS) {
            const double o = __shfl_xor_sync(0xffffffffu, acc[3], off);
            if (acc[3] > o) acc[3] = o;
          }
        }
        const double tmin = TS ? acc[3] : INF;
        if (SPLIT > 32) {
          __syncthreads();
          if (lane == 0) {
            red[4 * (tid >> 5)] = acc[0];
            red[4 * (tid >> 5) + 1] = acc[1];
            red[4 * (tid >> 5) + 2] = acc[2];
            red[4 * (tid >> 5) + 3] = tmin;
          }
          __syncthreads();
          if (tid < nac) {
            const int wpt = SPLIT >> 5;
            double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = INF;
            for (int w = 0; w < wpt; ++w) {
              q0 += red[4 * (tid * wpt + w)];
              q1 += red[4 * (tid * wpt + w) + 1];
              q2 += red[4 * (tid * wpt + w) + 2];
              const double o = red[4 * (tid * wpt + w) + 3];
              if (q3 > o) q3 = o;
            }
            double *r = res + 4 * (size_t)(rank + NC * (u0 + tid));
            r[0] = q0;
            r[1] = q1;
            r[2] = q2;
            r[3] = q3;
          }
        } else if (live && sl == 0) {
          double *r = res + 4 * (size_t)x;
          r[0] = acc[0];
          r[1] = acc[1];
          r[2] = acc[2];
          r[3] = tmin;
        }
      }
      __syncthreads();
      // v += dv; dt_max = min(dt_max, tmin) (leapfrog.ml:103-104); second half drift of the block
      for (int i = rank + NC * tid; i < iend; i += NC * T) {
        double *b = rows + (size_t)ms[i] * PK;
        const double *r = res + 4 * (size_t)i;
        b[4] = __dadd_rn(b[4], dt * r[0]);
        b[5] = __dadd_rn(b[5], dt * r[1]);
        b[6] = __dadd_rn(b[6], dt * r[2]);
        if (TS) {
          double d = dtm[i];  // an active slot holds infinity since stage 0
          if (d > r[3]) d = r[3];
          dtm[i] = d;
        }
        if (i >= ibegin) {
          tt[i] = __dadd_rn(tt[i], dt2);
          b[0] = __dadd_rn(b[0], __dmul_rn(dt2, b[4]));
          b[1] = __dadd_rn(b[1], __dmul_rn(dt2, b[5]));
          b[2] = __dadd_rn(b[2], __dmul_rn(dt2, b[6]));
        }
      }
      if constexpr (NC > 1) {
        cluster.sync();  // every CTA is done reading the old rows
        // own slots -> the other replicas: {q, m, v} row, dt_max, t  (10 doubles per slot)
        for (int e = tid; e < nslots * 10 * (NC - 1); e += T) {
          const int peer0 = e % (NC - 1), w = (e / (NC - 1)) % 10, i = rank + NC * (e / (10 * (NC - 1)));
          const int peer = peer0 + (peer0 >= rank ? 1 : 0);
          double *src = w < 8 ? rows + (size_t)ms[i] * PK + w : (w == 8 ? dtm + i : tt + i);
          *cluster.map_shared_rank(src, peer) = *src;
        }
        cluster.sync();  // the replicas agree again
      }
      nk += 1;
    }
    if (sp >= LFT_DEPTH) {
      err = 1;
      break;
    }
    if (tid == 0) {
      stack[sp - 1].stage = 2;
      stack[sp] = LfFrame{ibegin, 0, 0, 0, dt2};
    }
    ++sp;
    __syncthreads();
  }
  __syncthreads();
  if constexpr (NC > 1) cluster.sync();  // nobody leaves while a peer may still store into its shared memory
  if (rank == 0) {
    for (int e = tid; e < top * PK; e += T) P.packed[e] = rows[(size_t)ms[e / PK] * PK + (e % PK)];
    for (int i = tid; i < top; i += T) {
      P.dt_max[i] = dtm[i];
      P.t[i] = tt[i];
      P.map[i] = ms[i];
    }
    if (tid == 0) {
      P.out[0] = nk;
      P.out[1] = err;
    }
  }
}

// ---- O(N) helpers ----------------------------------------------------------------------------
// Pack (m, q, p|v) into the 64-byte source layout; v = p / m is the IEEE division the
// reference performs per pair (base.ml:93-97), done once per body.
__global__ void pack_kernel(int n, const double *__restrict__ m, const double *__restrict__ q,
                            const double *__restrict__ vel, int is_velocity,
                            double *__restrict__ packed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double mi = m[i];
  double2 *o = reinterpret_cast<double2 *>(packed + (size_t)i * PK);
  double vx = 0.0, vy = 0.0, vz = 0.0;
  if (vel) {
    vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
    if (!is_velocity) {
      vx = __ddiv_rn(vx, mi);
      vy = __ddiv_rn(vy, mi);
      vz = __ddiv_rn(vz, mi);
    }
  }
  o[0] = make_double2(q[3 * i], q[3 * i + 1]);
  o[1] = make_double2(q[3 * i + 2], mi);
  o[2] = make_double2(vx, vy);
  o[3] = make_double2(vz, 0.0);
}

// ---- device-resident Omelyan_advancer.OA (omelyan_advancer.ml:44-104) ------------------------
// O(N) halves of the operators; every formula repeats the host mirror's (= the reference's)
// operation sequence without contraction, so a resident step is bit-identical to the
// host-buffer one.  x[3n] -= f * g[3n]   (b_operate's p_i -= f gsum_i, :44-56; c_operate's
// p_i -= pf gsum*_i, :83-90)
__global__ void oa_axpy_kernel(double *__restrict__ x, const double *__restrict__ g, size_t n3,
                               double f) {
  size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n3) x[a] = __dsub_rn(x[a], __dmul_rn(f, g[a]));
}
// a_operate (:58-65): q += p * (h / (2 m))
__global__ void oa_drift_kernel(int n, const double *__restrict__ m, double *__restrict__ q,
                                const double *__restrict__ p, double h) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double f = __ddiv_rn(h, __dmul_rn(2.0, m[i]));
  for (int k = 0; k < 3; ++k) q[3 * i + k] = __dadd_rn(q[3 * i + k], __dmul_rn(p[3 * i + k], f));
}
// c_operate's q*_i = q_i - (h^2 / (24 m_i)) gsum_i (:70-78)
__global__ void oa_qstar_kernel(int n, const double *__restrict__ m, const double *__restrict__ q,
                                const double *__restrict__ g, double h, double *__restrict__ qs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double f = __ddiv_rn(__dmul_rn(h, h), __dmul_rn(24.0, m[i]));
  for (int k = 0; k < 3; ++k) qs[3 * i + k] = __dsub_rn(q[3 * i + k], __dmul_rn(f, g[3 * i + k]));
}
// Analysis.bodies_to_el_phase_space (analysis.ml:904-919) from the per-body potential phi:
// el[4i..] = {E_i, E_i/m, |L_i|, |L_i|/m}, E_i = |p|^2/(2m) + phi_i, L = q x p.
__global__ void el_phase_space_kernel(int n, const double *__restrict__ m,
                                      const double *__restrict__ q, const double *__restrict__ p,
                                      const double *__restrict__ phi, double *__restrict__ el) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
  const double px = p[3 * i], py = p[3 * i + 1], pz = p[3 * i + 2];
  const double lx = __dsub_rn(__dmul_rn(qy, pz), __dmul_rn(qz, py));
  const double ly = __dsub_rn(__dmul_rn(qz, px), __dmul_rn(qx, pz));
  const double lz = __dsub_rn(__dmul_rn(qx, py), __dmul_rn(qy, px));
  const double lb = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(lx, lx), __dmul_rn(ly, ly)), __dmul_rn(lz, lz)));
  const double p2 = __dadd_rn(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)), __dmul_rn(pz, pz));
  const double ke = __ddiv_rn(p2, __dmul_rn(2.0, m[i]));
  const double e = __dadd_rn(ke, phi[i]);
  el[4 * i + 0] = e;
  el[4 * i + 1] = __ddiv_rn(e, m[i]);
  el[4 * i + 2] = lb;
  el[4 * i + 3] = __ddiv_rn(lb, m[i]);
}

// Tangent direction (dq, dp|dv) -> {dq, 0, dv, 0}, dv = dp / m.
__global__ void pack_aux_kernel(int n, const double *__restrict__ m, const double *__restrict__ dq,
                                const double *__restrict__ dvel, int is_velocity,
                                double *__restrict__ aux) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2 *o = reinterpret_cast<double2 *>(aux + (size_t)i * PK);
  double vx = 0.0, vy = 0.0, vz = 0.0;
  if (dvel) {
    vx = dvel[3 * i], vy = dvel[3 * i + 1], vz = dvel[3 * i + 2];
    if (!is_velocity) {
      double mi = m[i];
      vx = __ddiv_rn(vx, mi);
      vy = __ddiv_rn(vy, mi);
      vz = __ddiv_rn(vz, mi);
    }
  }
  o[0] = make_double2(dq[3 * i], dq[3 * i + 1]);
  o[1] = make_double2(dq[3 * i + 2], 0.0);
  o[2] = make_double2(vx, vy);
  o[3] = make_double2(vz, 0.0);
}

// Hermite predictor fused with the pack (hermite.ml:53-72): body i is predicted from its own
// (t, q, p, f, df) to time nt with the reference's exact operation sequence (no contraction),
//   qp = q + pc p + fc f + dfc df,  pc = dt/m, fc = pc dt 0.5, dfc = fc dt / 3
//   pp = p + dt f + (dt dt 0.5) df
// and stored as {qp, m, pp/m}.
__global__ void predict_pack_kernel(int n, const double *__restrict__ t,
                                    const double *__restrict__ m, const double *__restrict__ q,
                                    const double *__restrict__ p, const double *__restrict__ f,
                                    const double *__restrict__ df, double nt,
                                    double *__restrict__ packed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mi = m[i];
  const double dt = __dsub_rn(nt, t[i]);
  const double pc = __ddiv_rn(dt, mi);
  const double fc = __dmul_rn(__dmul_rn(pc, dt), 0.5);
  const double dfc = __ddiv_rn(__dmul_rn(fc, dt), 3.0);
  const double pdfc = __dmul_rn(__dmul_rn(dt, dt), 0.5);
  double qp[3], vp[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double qk = q[3 * i + k], pk = p[3 * i + k], fk = f[3 * i + k], dfk = df[3 * i + k];
    qp[k] = __dadd_rn(__dadd_rn(__dadd_rn(qk, __dmul_rn(pc, pk)), __dmul_rn(fc, fk)),
                      __dmul_rn(dfc, dfk));
    const double pp = __dadd_rn(__dadd_rn(pk, __dmul_rn(dt, fk)), __dmul_rn(pdfc, dfk));
    vp[k] = __ddiv_rn(pp, mi);
  }
  double2 *o = reinterpret_cast<double2 *>(packed + (size_t)i * PK);
  o[0] = make_double2(qp[0], qp[1]);
  o[1] = make_double2(qp[2], mi);
  o[2] = make_double2(vp[0], vp[1]);
  o[3] = make_double2(vp[2], 0.0);
}

// The same predictor over dual numbers (dFloat_hermite.ml:60-79): direction k of body i carries
// tangents of t, q, p, f, df (the mass is a constant, dFloat_hermite.ml:36-46) and the target
// time nt carries nt_tan[k].  Values follow predict_pack_kernel; the tangents are the product /
// quotient rules applied to the same expression tree:
//   ddt = dnt - dt_i, dpc = ddt/m, dfc = (dpc dt + pc ddt) 0.5, ddfc = (dfc dt + fc ddt)/3
//   dqp = dq + dpc p + pc dp + dfc f + fc df_ + ddfc df + dfc_ ddf      (df_ = f's tangent)
//   dpp = dp + ddt f + dt df_ + (dt ddt) df + (dt dt 0.5) ddf
// stored as the tangent sweep's direction rows {dqp, 0, dpp/m, 0}, direction k at
// aux + k * aux_stride.  grid = (bodies, K).  t_tan / nt_tan may be NULL (zeros).
__global__ void predict_pack_tangent_kernel(
    int n, const double *__restrict__ t, const double *__restrict__ m, const double *__restrict__ q,
    const double *__restrict__ p, const double *__restrict__ f, const double *__restrict__ df,
    double nt, const double *__restrict__ t_tan, const double *__restrict__ q_tan,
    const double *__restrict__ p_tan, const double *__restrict__ f_tan,
    const double *__restrict__ df_tan, const double *__restrict__ nt_tan,
    double *__restrict__ aux, size_t aux_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (i >= n) return;
  const size_t o3 = (size_t)k * 3 * n + 3 * (size_t)i;
  const double mi = m[i];
  const double dt = __dsub_rn(nt, t[i]);
  const double ddt = (nt_tan ? nt_tan[k] : 0.0) - (t_tan ? t_tan[(size_t)k * n + i] : 0.0);
  const double pc = __ddiv_rn(dt, mi), dpc = ddt / mi;
  const double fc = __dmul_rn(__dmul_rn(pc, dt), 0.5), dfc = (dpc * dt + pc * ddt) * 0.5;
  const double dfc3 = __ddiv_rn(__dmul_rn(fc, dt), 3.0), ddfc3 = (dfc * dt + fc * ddt) / 3.0;
  const double pdfc = __dmul_rn(__dmul_rn(dt, dt), 0.5), dpdfc = dt * ddt;
  double dqp[3], dvp[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double pk = p[3 * i + c], fk = f[3 * i + c], dfk = df[3 * i + c];
    const double dq = q_tan[o3 + c], dp = p_tan[o3 + c], dff = f_tan[o3 + c], ddf = df_tan[o3 + c];
    dqp[c] = dq + (dpc * pk + pc * dp) + (dfc * fk + fc * dff) + (ddfc3 * dfk + dfc3 * ddf);
    const double dpp = dp + (ddt * fk + dt * dff) + (dpdfc * dfk + pdfc * ddf);
    dvp[c] = dpp / mi;
  }
  double2 *o = reinterpret_cast<double2 *>(aux + (size_t)k * aux_stride + (size_t)i * PK);
  o[0] = make_double2(dqp[0], dqp[1]);
  o[1] = make_double2(dqp[2], 0.0);
  o[2] = make_double2(dvp[0], dvp[1]);
  o[3] = make_double2(dvp[2], 0.0);
}

}  // namespace nbk
