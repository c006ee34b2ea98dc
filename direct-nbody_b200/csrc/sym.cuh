// Pair-once (symmetric) Hermite acc+jerk sweep for sm_100a.
//
// The reference's own start_bs loop (hermite.ml:157-166) evaluates every unordered pair once and
// applies it to both bodies (f_i -= gv; f_j += gv).  sweep_kernel (sweep.cuh) evaluates every
// ORDERED pair: 31 FP64 instructions per interaction.  Here a pair evaluation costs 38 FP64
// instructions and serves two interactions = 19 per interaction.
//
// Decomposition: the bodies are cut into blocks of SYM_B = 1024 rows; a work item is one block
// pair (I, J), I <= J, of the upper triangle.  A CTA of 256 threads holds 4 targets of block I
// per thread in registers (position, velocity, mass, six accumulators) and block J in shared
// memory (three double2 arrays and one double array, see SYM_SMEM) plus six j-side accumulator
// arrays.  A warp works on one group of 32 sources at a time: in round r lane l takes source
// (l + r) mod 32 of the group (every lane a different one: bank-conflict-free), evaluates it
// against its 4 targets, adds the i-side terms
// to the target accumulators and the j-side terms (weighted with the TARGET's mass) to six
// registers that travel with the source: after every round they move one lane down by SHFL, so
// after 32 rounds every lane holds the complete sum over the warp's 128 targets for one source
// and adds it to the shared arrays.  The 8 warps walk the 32 groups skewed (warp w starts at
// group 4w), so no two warps ever touch the same group between two CTA barriers (one barrier
// per 4 groups).  At the end of an item both sides leave as partial sums: body b of block B has
// one slot per partner block S, written by item (min(B,S), max(B,S)); sym_reduce_kernel adds a
// body's slots in slot order, so the result does not depend on which CTA (or which GPU) ran
// which item.  Diagonal items (I == J) evaluate all ordered pairs of the block one-sidedly
// (j-side weight 0) and mask i == j; items touching the padded tail of the last block mask the
// rows >= n.
//
// Sharded runs: body b lives in row b % shard_rows of shard[b / shard_rows]; every rank runs a
// contiguous range of items behind the owners' ready flags and the per-rank sums meet in
// sym_gather_kernel (each rank's peer reads of its slice) or an NCCL reduce-scatter
// (nbk/sharded.py).  The shards may be peer-mapped rows on their owners' GPUs or - what
// PeerShardedAccJerkSym does - pieces of a local array every rank's pack_push_kernel has
// stored its rows into.
//
// sym_lite_kernel (end of this file): the same decomposition for the sums without velocities,
// Base.grad_v and Base.v.
//
// SYM_UNROLL / SYM_NEWTON: A/B knobs of tools/time_sym.py builds (rounds unrolled in the hot
// loop; association of the last rsqrt step); the defaults are the measured best.
#pragma once

namespace nbk {

#ifndef SYM_UNROLL
#define SYM_UNROLL 2
#endif
#ifndef SYM_NEWTON
#define SYM_NEWTON 1
#endif
constexpr int SYM_ROUNDS_UNROLLED = SYM_UNROLL;
constexpr int SYM_NT = 256;
constexpr int SYM_IPT = 4;
constexpr int SYM_B = SYM_NT * SYM_IPT;                       // 1024 bodies per block
// Shared memory: the sources as three double2 arrays {x,y} {z,m} {vx,vy} + vz (3 LDS.128 +
// 1 LDS.64 per source), every group of 32 stored TWICE in a row (64 entries) so that lane l
// reads entry l + round without wrapping the index, + six arrays of j-side sums: 160 KB.
constexpr size_t SYM_SMEM = (size_t)(2 * 7 + 6) * SYM_B * sizeof(double);

struct SymParams {
  const double *shard[MAX_PEERS];  // packed rows [shard_rows][PK] of every owner
  long long shard_rows;            // multiple of SYM_B (single device: >= nb * SYM_B)
  int n, nb, nranks, self_rank;
  long long item0, item1;          // items [item0, item1) of the upper triangle, row-major
  double eps2;
  double *partial;                 // [nb slots][nb * SYM_B bodies][6]
  const unsigned long long *ready;
  unsigned long long epoch, wait_ns;
};

__host__ __device__ inline long long sym_item_of(int I, int J, int nb) {
  return (long long)I * nb - (long long)I * (I - 1) / 2 + (J - I);
}
__device__ inline void sym_decode(long long k, int nb, int &I, int &J) {
  const double b = 2.0 * nb + 1.0;
  int i = (int)((b - sqrt(b * b - 8.0 * (double)k)) * 0.5);
  i = max(0, min(nb - 1, i));
  while (i > 0 && sym_item_of(i, i, nb) > k) --i;
  while (i + 1 < nb && sym_item_of(i + 1, i + 1, nb) <= k) ++i;
  I = i;
  J = i + (int)(k - sym_item_of(i, i, nb));
}

__device__ __forceinline__ const double *sym_block_rows(const SymParams &P, int blk) {
  const long long b0 = (long long)blk * SYM_B;
  const int owner = (int)(b0 / P.shard_rows);
  return P.shard[owner] + (size_t)(b0 - (long long)owner * P.shard_rows) * PK;
}

// One item.  CHECK: mask self pairs (DIAG) and rows >= n; DIAG also switches the j side off.
template <bool CHECK>
__device__ __forceinline__ void sym_item(const SymParams &P, const int I, const int J, double *sm) {
  constexpr int B = SYM_B;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double2 *sxy = reinterpret_cast<double2 *>(sm), *szm = sxy + 2 * B, *svxy = sxy + 4 * B;
  double *svz = sm + 12 * B, *jacc = sm + 14 * B;
  const bool diag = I == J;
  const double eps2 = P.eps2;

  double tx[SYM_IPT], ty[SYM_IPT], tz[SYM_IPT], tvx[SYM_IPT], tvy[SYM_IPT], tvz[SYM_IPT],
      tm[SYM_IPT];
  int ig[SYM_IPT];
  {
    const double2 *rows = reinterpret_cast<const double2 *>(sym_block_rows(P, I));
#pragma unroll
    for (int r = 0; r < SYM_IPT; ++r) {
      const int il = r * SYM_NT + tid;
      ig[r] = I * B + il;
      double2 a = make_double2(0.0, 0.0), b = a, c = a, d = a;
      if (!CHECK || ig[r] < P.n) {
        a = rows[(size_t)il * 4 + 0];
        b = rows[(size_t)il * 4 + 1];
        c = rows[(size_t)il * 4 + 2];
        d = rows[(size_t)il * 4 + 3];
      }
      tx[r] = a.x, ty[r] = a.y, tz[r] = b.x, tvx[r] = c.x, tvy[r] = c.y, tvz[r] = d.x;
      // the j side is weighted with the target's mass; diagonal items leave it out
      tm[r] = (CHECK && (diag || ig[r] >= P.n)) ? 0.0 : b.y;
    }
  }
  {
    const double2 *rows = reinterpret_cast<const double2 *>(sym_block_rows(P, J));
#pragma unroll
    for (int r = 0; r < SYM_IPT; ++r) {
      const int jl = r * SYM_NT + tid;
      double2 a = make_double2(0.0, 0.0), b = a, c = a, d = a;
      if (!CHECK || J * B + jl < P.n) {
        a = rows[(size_t)jl * 4 + 0];
        b = rows[(size_t)jl * 4 + 1];
        c = rows[(size_t)jl * 4 + 2];
        d = rows[(size_t)jl * 4 + 3];
      }
      const int e = ((jl >> 5) << 6) + (jl & 31);  // entry and its copy 32 further on
      sxy[e] = a, szm[e] = b, svxy[e] = c, svz[e] = d.x;
      sxy[e + 32] = a, szm[e + 32] = b, svxy[e + 32] = c, svz[e + 32] = d.x;
#pragma unroll
      for (int c6 = 0; c6 < 6; ++c6) jacc[c6 * B + jl] = 0.0;
    }
  }
  double acc[SYM_IPT][6];
#pragma unroll
  for (int r = 0; r < SYM_IPT; ++r)
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[r][k] = 0.0;
  __syncthreads();

  const int jbase = J * B;
  const int src_lane = (lane + 1) & 31;
  for (int g = 0; g < 8; ++g) {
    for (int s4 = 0; s4 < 4; ++s4) {
      const int grp = ((((warp + g) & 7) << 2) + s4) << 5;  // first source of the group
      const int gent = 2 * grp + lane;
      // one round: the lane's source of this round against its 4 targets
      auto round = [&](const int rd, double (&ja)[6]) {
        const int e = gent + rd;  // source grp + (lane + rd) mod 32, through the doubled group
        const double2 jxy = sxy[e], jzm = szm[e], jvxy = svxy[e];
        const double jx = jxy.x, jy = jxy.y, jz = jzm.x, jm = jzm.y;
        const double jvx = jvxy.x, jvy = jvxy.y, jvz = svz[e];
        double dx[SYM_IPT], dy[SYM_IPT], dz[SYM_IPT], ux[SYM_IPT], uy[SYM_IPT], uz[SYM_IPT];
        double r2[SYM_IPT], rv[SYM_IPT], y0[SYM_IPT], y[SYM_IPT], wj[SYM_IPT], wi[SYM_IPT],
            a3[SYM_IPT];
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r) {
          dx[r] = jx - tx[r];
          dy[r] = jy - ty[r];
          dz[r] = jz - tz[r];
        }
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r) {
          r2[r] = fma(dx[r], dx[r], eps2);
          r2[r] = fma(dy[r], dy[r], r2[r]);
          r2[r] = fma(dz[r], dz[r], r2[r]);
          wj[r] = jm;
          wi[r] = tm[r];
          if (CHECK) {
            const int jg = jbase + grp + ((lane + rd) & 31);
            const bool off = (jg == ig[r]) || (jg >= P.n) || (ig[r] >= P.n);
            r2[r] = off ? 1.0 : r2[r];
            wj[r] = off ? 0.0 : wj[r];
            wi[r] = off ? 0.0 : wi[r];
          }
        }
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r)
          asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[r]) : "d"(r2[r]));
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r) {
          ux[r] = jvx - tvx[r];
          uy[r] = jvy - tvy[r];
          uz[r] = jvz - tvz[r];
        }
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r) {
          rv[r] = dx[r] * ux[r];
          rv[r] = fma(dy[r], uy[r], rv[r]);
          rv[r] = fma(dz[r], uz[r], rv[r]);
        }
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r) {
          const double tt = r2[r] * y0[r];
          const double e = fma(-tt, y0[r], 1.0);
          const double pp = fma(0.375, e, 0.5);
#if SYM_NEWTON
          const double q = e * pp;
          y[r] = fma(y0[r], q, y0[r]);
#else
          const double q = e * y0[r];
          y[r] = fma(pp, q, y0[r]);
#endif
        }
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r) {
          const double y2 = y[r] * y[r];
          const double y3 = y[r] * y2;
          a3[r] = (rv[r] * y2) * -3.0;
          wj[r] *= y3;
          wi[r] *= y3;
        }
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r) {
          acc[r][0] = fma(wj[r], dx[r], acc[r][0]);
          acc[r][1] = fma(wj[r], dy[r], acc[r][1]);
          acc[r][2] = fma(wj[r], dz[r], acc[r][2]);
          ja[0] = fma(wi[r], dx[r], ja[0]);
          ja[1] = fma(wi[r], dy[r], ja[1]);
          ja[2] = fma(wi[r], dz[r], ja[2]);
        }
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r) {
          ux[r] = fma(a3[r], dx[r], ux[r]);
          uy[r] = fma(a3[r], dy[r], uy[r]);
          uz[r] = fma(a3[r], dz[r], uz[r]);
        }
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r) {
          acc[r][3] = fma(wj[r], ux[r], acc[r][3]);
          acc[r][4] = fma(wj[r], uy[r], acc[r][4]);
          acc[r][5] = fma(wj[r], uz[r], acc[r][5]);
          ja[3] = fma(wi[r], ux[r], ja[3]);
          ja[4] = fma(wi[r], uy[r], ja[4]);
          ja[5] = fma(wi[r], uz[r], ja[5]);
        }
      };
      double ja[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) ja[k] = 0.0;
#pragma unroll SYM_ROUNDS_UNROLLED
      for (int rd = 0; rd < 32; ++rd) {
        round(rd, ja);
        // the source's sums follow it to the lane that takes it next round
#pragma unroll
        for (int k = 0; k < 6; ++k) ja[k] = __shfl_sync(0xffffffffu, ja[k], src_lane);
      }
      // 32 moves = a full turn: lane l holds the sums of source grp + l
#pragma unroll
      for (int k = 0; k < 6; ++k) jacc[k * B + grp + lane] += ja[k];
    }
    __syncthreads();
  }

  // Both sides leave as partial sums: slot J of block I (targets), slot I of block J (sources,
  // with the sign of the reversed separation).
  const size_t nbB = (size_t)P.nb * B;
#pragma unroll
  for (int r = 0; r < SYM_IPT; ++r) {
    double2 *dst = reinterpret_cast<double2 *>(P.partial + ((size_t)J * nbB + (size_t)ig[r]) * 6);
    dst[0] = make_double2(acc[r][0], acc[r][1]);
    dst[1] = make_double2(acc[r][2], acc[r][3]);
    dst[2] = make_double2(acc[r][4], acc[r][5]);
  }
  if (!diag) {
#pragma unroll
    for (int r = 0; r < SYM_IPT; ++r) {
      const int jl = r * SYM_NT + tid;
      double2 *dst =
          reinterpret_cast<double2 *>(P.partial + ((size_t)I * nbB + (size_t)(jbase + jl)) * 6);
      dst[0] = make_double2(-jacc[jl], -jacc[B + jl]);
      dst[1] = make_double2(-jacc[2 * B + jl], -jacc[3 * B + jl]);
      dst[2] = make_double2(-jacc[4 * B + jl], -jacc[5 * B + jl]);
    }
  }
  __syncthreads();  // the shared arrays are refilled by the next item
}

__global__ void __launch_bounds__(SYM_NT, 1) sym_sweep_kernel(const SymParams P) {
  extern __shared__ __align__(16) unsigned char sym_smem_raw[];
  double *sm = reinterpret_cast<double *>(sym_smem_raw);
  if (P.ready != nullptr) {
    if (threadIdx.x == 0) {
      unsigned long long t_start;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
      for (int r = 0; r < P.nranks; ++r)
        if (r != P.self_rank)
          while (ld_acquire_sys(P.ready + r) < P.epoch) {
            __nanosleep(64);
            unsigned long long t_now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
            if (P.wait_ns != 0 && t_now - t_start > P.wait_ns) __trap();
          }
    }
    __syncthreads();
  }
  const bool ragged = (long long)P.nb * SYM_B != (long long)P.n;
  for (long long k = P.item0 + blockIdx.x; k < P.item1; k += gridDim.x) {
    int I, J;
    sym_decode(k, P.nb, I, J);
    if (I == J || (ragged && J == P.nb - 1))
      sym_item<true>(P, I, J, sm);
    else
      sym_item<false>(P, I, J, sm);
  }
}

// out6[b][0..5] (or gsum / dgsum rows) = -m_b * sum over the slots items [item0, item1) filled,
// in slot order.  One thread per (body, component).
struct SymReduceParams {
  const double *shard[MAX_PEERS];
  long long shard_rows;
  int n, nb;
  long long item0, item1;
  const double *partial;
  double *out6;    // [n][6] or NULL
  double *gsum;    // [t1 - t0][3] (with dgsum) or NULL
  double *dgsum;
  int t0, t1;      // bodies written
};

__global__ void __launch_bounds__(192) sym_reduce_kernel(const SymReduceParams R) {
  // 32 bodies x 6 components per CTA
  const int b = R.t0 + blockIdx.x * 32 + threadIdx.x / 6;
  const int c = threadIdx.x % 6;
  if (b >= R.t1) return;
  const int Bk = b / SYM_B;
  const size_t nbB = (size_t)R.nb * SYM_B;
  const double *src = R.partial + (size_t)b * 6 + c;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const bool all = R.item0 == 0 && R.item1 == sym_item_of(R.nb - 1, R.nb - 1, R.nb) + 1;
  if (all) {
    int s = 0;
    for (; s + 4 <= R.nb; s += 4) {
      s0 += src[(size_t)s * nbB * 6];
      s1 += src[(size_t)(s + 1) * nbB * 6];
      s2 += src[(size_t)(s + 2) * nbB * 6];
      s3 += src[(size_t)(s + 3) * nbB * 6];
    }
    for (; s < R.nb; ++s) s0 += src[(size_t)s * nbB * 6];
  } else {
    for (int s = 0; s < R.nb; ++s) {
      const long long k = sym_item_of(min(Bk, s), max(Bk, s), R.nb);
      if (k >= R.item0 && k < R.item1) s0 += src[(size_t)s * nbB * 6];
    }
  }
  const long long owner = (long long)b / R.shard_rows;
  const double m = R.shard[owner][(size_t)((long long)b - owner * R.shard_rows) * PK + 3];
  const double v = -m * ((s0 + s1) + (s2 + s3));
  if (R.out6) R.out6[(size_t)b * 6 + c] = v;
  if (R.gsum) {
    if (c < 3) R.gsum[(size_t)(b - R.t0) * 3 + c] = v;
    else R.dgsum[(size_t)(b - R.t0) * 3 + (c - 3)] = v;
  }
}

// Bodies [a, b): the per-rank sums sum6[e][body][0..5] (peer-mapped) added in rank order ->
// gsum, dgsum rows [b - a][3]: the reduce-scatter of the sharded pair-once sweep as the ranks'
// own peer reads.  ready != NULL: rank e's sums may be read once ready[e] >= epoch (the rank
// publishes that flag array after its slot reduction, nbk_dev_publish); NULL: the caller
// ordered the streams (nbk_create_multi).
struct SymGatherParams {
  const double *sum6[MAX_PEERS];
  int ndev, self, a, b;
  double *g, *dg;
  const unsigned long long *ready;
  unsigned long long epoch, wait_ns;
};
__global__ void sym_gather_kernel(const SymGatherParams G) {
  if (G.ready != nullptr) {
    if (threadIdx.x == 0) {
      unsigned long long t_start;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
      for (int e = 0; e < G.ndev; ++e)
        if (e != G.self)
          while (ld_acquire_sys(G.ready + e) < G.epoch) {
            __nanosleep(64);
            unsigned long long t_now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
            if (G.wait_ns != 0 && t_now - t_start > G.wait_ns) __trap();
          }
    }
    __syncthreads();
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (G.b - G.a) * 6) return;
  const int body = i / 6, c = i % 6;
  double s = 0.0;
  for (int e = 0; e < G.ndev; ++e) s += G.sum6[e][(size_t)(G.a + body) * 6 + c];
  if (c < 3) G.g[body * 3 + c] = s;
  else G.dg[body * 3 + c - 3] = s;
}

// Pack (m, q, p|v) rows as pack_kernel does and store every row into nranks destinations: this
// rank's place in every rank's gathered row array (peer-mapped over NVLink) - the all-gather of
// the sharded pair-once sweep as remote stores of the packing kernel itself.  The publish
// kernel that follows on the stream releases the flag at system scope.
struct PackPushParams {
  double *dst[MAX_PEERS];
};
__global__ void pack_push_kernel(int n, const double *__restrict__ m, const double *__restrict__ q,
                                 const double *__restrict__ vel, int is_velocity, int nranks,
                                 const PackPushParams D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double mi = m[i];
    double vx = 0.0, vy = 0.0, vz = 0.0;
    if (vel) {
      vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
      if (!is_velocity) {
        vx = __ddiv_rn(vx, mi);
        vy = __ddiv_rn(vy, mi);
        vz = __ddiv_rn(vz, mi);
      }
    }
    const double2 a = make_double2(q[3 * i], q[3 * i + 1]), b = make_double2(q[3 * i + 2], mi);
    const double2 c = make_double2(vx, vy), d = make_double2(vz, 0.0);
    for (int r = 0; r < nranks; ++r) {
      double2 *o = reinterpret_cast<double2 *>(D.dst[r] + (size_t)i * PK);
      o[0] = a, o[1] = b, o[2] = c, o[3] = d;
    }
  }
  __threadfence_system();
}

// A rank's slice of the reduce-scattered sums [count][6] -> gsum, dgsum [count][3].
__global__ void sym_split_kernel(const double *__restrict__ in6, int count, double *__restrict__ g,
                                 double *__restrict__ dg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count * 6) return;
  const int b = i / 6, c = i % 6;
  if (c < 3) g[b * 3 + c] = in6[i];
  else dg[b * 3 + c - 3] = in6[i];
}

// ---- pair-once sweeps without velocities: Base.grad_v (OP = SYM_ACC) and Base.v (SYM_POT) --------
// The same decomposition for the acceleration-only sum (b_operate / c_operate of the Omelyan
// advancer, omelyan_advancer.ml:44-94; 21 FP64 instructions per pair evaluation = 10.5 per
// interaction against 17 in the ordered sweep) and the potential (energy.ml:66-73,
// analysis.ml:904-919; 13 = 6.5 per interaction against 12).  One device, all items.
enum { SYM_ACC = 1, SYM_POT = 2 };
template <int OP> struct SymLite {
  static constexpr int NACC = OP == SYM_ACC ? 3 : 1;
};
constexpr size_t SYM_LITE_SMEM = (size_t)(2 * 4 + 3) * SYM_B * sizeof(double);

struct SymLiteParams {
  const double *rows;   // packed rows of the n bodies
  int n, nb;
  long long items;
  double eps2;
  double *partial;      // [nb slots][nb * SYM_B bodies][NACC]
  double *out;          // [n][NACC], scaled with -m_b by the reduction
};

template <bool CHECK, int OP>
__device__ __forceinline__ void sym_lite_item(const SymLiteParams &P, const int I, const int J,
                                              double *sm) {
  constexpr int B = SYM_B, NA = SymLite<OP>::NACC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double2 *sxy = reinterpret_cast<double2 *>(sm), *szm = sxy + 2 * B;
  double *jacc = sm + 8 * B;
  const bool diag = I == J;
  const double eps2 = P.eps2;
  double tx[SYM_IPT], ty[SYM_IPT], tz[SYM_IPT], tm[SYM_IPT];
  int ig[SYM_IPT];
  {
    const double2 *rows = reinterpret_cast<const double2 *>(P.rows + (size_t)I * B * PK);
#pragma unroll
    for (int r = 0; r < SYM_IPT; ++r) {
      const int il = r * SYM_NT + tid;
      ig[r] = I * B + il;
      double2 a = make_double2(0.0, 0.0), b = a;
      if (!CHECK || ig[r] < P.n) {
        a = rows[(size_t)il * 4 + 0];
        b = rows[(size_t)il * 4 + 1];
      }
      tx[r] = a.x, ty[r] = a.y, tz[r] = b.x;
      tm[r] = (CHECK && (diag || ig[r] >= P.n)) ? 0.0 : b.y;
    }
  }
  {
    const double2 *rows = reinterpret_cast<const double2 *>(P.rows + (size_t)J * B * PK);
#pragma unroll
    for (int r = 0; r < SYM_IPT; ++r) {
      const int jl = r * SYM_NT + tid;
      double2 a = make_double2(0.0, 0.0), b = a;
      if (!CHECK || J * B + jl < P.n) {
        a = rows[(size_t)jl * 4 + 0];
        b = rows[(size_t)jl * 4 + 1];
      }
      const int e = ((jl >> 5) << 6) + (jl & 31);
      sxy[e] = a, szm[e] = b;
      sxy[e + 32] = a, szm[e + 32] = b;
#pragma unroll
      for (int c = 0; c < NA; ++c) jacc[c * B + jl] = 0.0;
    }
  }
  double acc[SYM_IPT][NA];
#pragma unroll
  for (int r = 0; r < SYM_IPT; ++r)
#pragma unroll
    for (int k = 0; k < NA; ++k) acc[r][k] = 0.0;
  __syncthreads();

  const int jbase = J * B;
  const int src_lane = (lane + 1) & 31;
  for (int g = 0; g < 8; ++g) {
    for (int s4 = 0; s4 < 4; ++s4) {
      const int grp = ((((warp + g) & 7) << 2) + s4) << 5;
      const int gent = 2 * grp + lane;
      double ja[NA];
#pragma unroll
      for (int k = 0; k < NA; ++k) ja[k] = 0.0;
#pragma unroll 2
      for (int rd = 0; rd < 32; ++rd) {
        const int e = gent + rd;
        const double2 jxy = sxy[e], jzm = szm[e];
        double dx[SYM_IPT], dy[SYM_IPT], dz[SYM_IPT], r2[SYM_IPT], y0[SYM_IPT], y[SYM_IPT],
            wj[SYM_IPT], wi[SYM_IPT];
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r) {
          dx[r] = jxy.x - tx[r];
          dy[r] = jxy.y - ty[r];
          dz[r] = jzm.x - tz[r];
        }
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r) {
          r2[r] = fma(dx[r], dx[r], eps2);
          r2[r] = fma(dy[r], dy[r], r2[r]);
          r2[r] = fma(dz[r], dz[r], r2[r]);
          wj[r] = jzm.y;
          wi[r] = tm[r];
          if (CHECK) {
            const int jg = jbase + grp + ((lane + rd) & 31);
            const bool off = (jg == ig[r]) || (jg >= P.n) || (ig[r] >= P.n);
            r2[r] = off ? 1.0 : r2[r];
            wj[r] = off ? 0.0 : wj[r];
            wi[r] = off ? 0.0 : wi[r];
          }
        }
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r)
          asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[r]) : "d"(r2[r]));
#pragma unroll
        for (int r = 0; r < SYM_IPT; ++r) {
          const double tt = r2[r] * y0[r];
          const double ee = fma(-tt, y0[r], 1.0);
          const double pp = fma(0.375, ee, 0.5);
          const double q = ee * pp;
          y[r] = fma(y0[r], q, y0[r]);
        }
        if constexpr (OP == SYM_ACC) {
#pragma unroll
          for (int r = 0; r < SYM_IPT; ++r) {
            const double y3 = y[r] * (y[r] * y[r]);
            wj[r] *= y3;
            wi[r] *= y3;
          }
#pragma unroll
          for (int r = 0; r < SYM_IPT; ++r) {
            acc[r][0] = fma(wj[r], dx[r], acc[r][0]);
            acc[r][1] = fma(wj[r], dy[r], acc[r][1]);
            acc[r][2] = fma(wj[r], dz[r], acc[r][2]);
            ja[0] = fma(wi[r], dx[r], ja[0]);
            ja[1] = fma(wi[r], dy[r], ja[1]);
            ja[2] = fma(wi[r], dz[r], ja[2]);
          }
        } else {
#pragma unroll
          for (int r = 0; r < SYM_IPT; ++r) {
            acc[r][0] = fma(wj[r], y[r], acc[r][0]);
            ja[0] = fma(wi[r], y[r], ja[0]);
          }
        }
#pragma unroll
        for (int k = 0; k < NA; ++k) ja[k] = __shfl_sync(0xffffffffu, ja[k], src_lane);
      }
#pragma unroll
      for (int k = 0; k < NA; ++k) jacc[k * B + grp + lane] += ja[k];
    }
    __syncthreads();
  }

  // slot J of block I (targets), slot I of block J (sources; the acceleration changes sign with
  // the separation, the potential does not)
  const size_t nbB = (size_t)P.nb * B;
  constexpr double JS = OP == SYM_ACC ? -1.0 : 1.0;
#pragma unroll
  for (int r = 0; r < SYM_IPT; ++r) {
    double *dst = P.partial + ((size_t)J * nbB + (size_t)ig[r]) * NA;
#pragma unroll
    for (int k = 0; k < NA; ++k) dst[k] = acc[r][k];
  }
  if (!diag) {
#pragma unroll
    for (int r = 0; r < SYM_IPT; ++r) {
      const int jl = r * SYM_NT + tid;
      double *dst = P.partial + ((size_t)I * nbB + (size_t)(jbase + jl)) * NA;
#pragma unroll
      for (int k = 0; k < NA; ++k) dst[k] = JS * jacc[k * B + jl];
    }
  }
  __syncthreads();
}

template <int OP>
__global__ void __launch_bounds__(SYM_NT, 1) sym_lite_kernel(const SymLiteParams P) {
  extern __shared__ __align__(16) unsigned char sym_smem_raw[];
  double *sm = reinterpret_cast<double *>(sym_smem_raw);
  const bool ragged = (long long)P.nb * SYM_B != (long long)P.n;
  for (long long k = blockIdx.x; k < P.items; k += gridDim.x) {
    int I, J;
    sym_decode(k, P.nb, I, J);
    if (I == J || (ragged && J == P.nb - 1))
      sym_lite_item<true, OP>(P, I, J, sm);
    else
      sym_lite_item<false, OP>(P, I, J, sm);
  }
}

// out[b][c] = -m_b * sum of body b's nb slots, in slot order.  One thread per (body, component).
template <int NA>
__global__ void __launch_bounds__(192) sym_lite_reduce_kernel(const SymLiteParams P) {
  constexpr int PER = 192 / NA;
  const int b = blockIdx.x * PER + threadIdx.x / NA;
  const int c = threadIdx.x % NA;
  if (b >= P.n) return;
  const size_t nbB = (size_t)P.nb * SYM_B;
  const double *src = P.partial + (size_t)b * NA + c;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int s = 0;
  for (; s + 4 <= P.nb; s += 4) {
    s0 += src[(size_t)s * nbB * NA];
    s1 += src[(size_t)(s + 1) * nbB * NA];
    s2 += src[(size_t)(s + 2) * nbB * NA];
    s3 += src[(size_t)(s + 3) * nbB * NA];
  }
  for (; s < P.nb; ++s) s0 += src[(size_t)s * nbB * NA];
  const double m = P.rows[(size_t)b * PK + 3];
  P.out[(size_t)b * NA + c] = -m * ((s0 + s1) + (s2 + s3));
}

}  // namespace nbk
