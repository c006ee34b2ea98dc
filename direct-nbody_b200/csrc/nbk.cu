// nbk.cu — libnbk: the C ABI of include/nbk.h over the sweep engine of sweep.cuh.
// sm_100a only; no CPU fallback (nbk_create fails without a CUDA device).
#include "../../include/nbk.h"
#include "sweep.cuh"
#include "sym.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace nbk;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

// One host thread per extra device of a multi context (nbk_create_multi): the per-device halves
// of a sharded sweep (upload + pack, then waits + launch + read-back) are enqueued side by side
// instead of from one serial host loop - the sweeps cannot start before EVERY device has staged
// its rows, so that loop's length was on the critical path.  Threads are internal: they never
// call back into the host program and are parked on a condition variable between calls.
struct MultiPool {
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  unsigned long long generation = 0;
  int pending = 0;
  bool stop = false;
  std::function<void(int)> job;
  // sense-reversing barrier between the phases of a job (all ndev participants)
  std::atomic<int> arrived{0}, sense{0};
  int nparticipants = 1;
  void barrier() {
    const int s = sense.load(std::memory_order_acquire);
    if (arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == nparticipants) {
      arrived.store(0, std::memory_order_relaxed);
      sense.store(s ^ 1, std::memory_order_release);
    } else {
      while (sense.load(std::memory_order_acquire) == s) std::this_thread::yield();
    }
  }
};

}  // namespace

struct nbk_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_valid = false;
  int64_t launches = 0;
  int shape = 0;
  unsigned long long peer_wait_ns = 20000000000ULL;  // nbk_set_peer_timeout / NBK_PEER_TIMEOUT_S
  std::string err;
  // device workspaces
  DevBuf d_m, d_q, d_vel, d_t, d_f, d_df, d_aux_in, d_aux_in2, d_packed, d_aux, d_partial, d_tbest,
      d_out[4], d_scalar, d_sympartial;
  int sym_mode = 1;  // pair-once sweep for large full squares: 1 auto, 0 never, 2 whenever possible (NBK_SYM)
  // resident bodies
  int resident_n = 0;
  DevBuf d_res_m, d_res_packed;
  // resident Advancer.A records
  int adv_n = 0;
  DevBuf d_adv, d_adv2;
  DevBuf d_adt;  // workspace of the device-side advance_range (nbk_adv_advance_range)
  DevBuf d_lft;  // workspace of the device-side advance_subsystem (nbk_lf_advance_subsystem)
  bool lft_attr_set = false;
  bool adt_attr_set = false;
  long long adt_prof[11] = {0};  // block steps, 9 per-phase cycle counts of CTA 0, kernel launches
  // resident Leapfrog.b array: packed rows {q, m, v} in array order + dt_max
  int lf_n = 0;
  DevBuf d_lf_packed, d_lf_packed2, d_lf_dtmax, d_lf_dtmax2;
  double *h_lf_stage = nullptr;  // pinned staging: dt_max read-backs and sort permutations
  size_t h_lf_cap = 0;
  cudaEvent_t ev_lf = nullptr;   // the last permutation upload has left the staging buffer
  // resident Omelyan_advancer.OA state: m, q, p, q*, gsum; g_valid = gsum belongs to the current q
  int oa_n = 0;
  bool oa_g_valid = false;
  double oa_g_eps2 = 0.0;
  DevBuf d_oa_m, d_oa_q, d_oa_p, d_oa_qs, d_oa_g, d_oa_phi, d_oa_el;
  // resident Hermite state
  int hermite_n = 0;
  int hermite_threads = 0;  // 0 auto; 128/256/512/1024 force the stepping CTA size, -8 the cluster kernel (NBK_HERMITE_THREADS)
  bool hermite_spec = true; // single-CTA loop: software-pipelined (default) or plain (NBK_HERMITE_SPEC=0)
  DevBuf d_hstate, d_hpartial, d_hticket, d_hout, d_hgrid;
  double *h_hres = nullptr, *d_hres = nullptr;  // mapped pinned result slot of the per-step kernel
  double hseq = 0.0;
  // pinned host staging for scalar read-back
  double *h_scalar = nullptr;
  // single-process multi-GPU (nbk_create_multi): further contexts, one per extra device
  std::vector<nbk_ctx *> peers;
  cudaEvent_t ev_ready = nullptr;
  cudaEvent_t ev_sum = nullptr;  // multi context: this device's pair-once sums are complete
  bool p2p_all = false;      // every device of the multi context can map every other's memory
  bool multi_sharded = false;  // the last multi_stage left the bodies sharded (peer-read sweeps)
  int multi_shard_rows = 0;
  MultiPool *pool = nullptr;   // worker threads of a multi context (NULL: serial host loop)
};

namespace {

int fail(nbk_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  else g_create_error = buf;
  return code;
}

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(ctx, NBK_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                    \
  } while (0)

int reserve(nbk_ctx *ctx, DevBuf &b, size_t bytes) {
  if (bytes <= b.cap) return NBK_OK;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess)
    return fail(ctx, NBK_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  b.cap = want;
  return NBK_OK;
}

#define TRY(expr)               \
  do {                          \
    int rc_ = (expr);           \
    if (rc_ != NBK_OK) return rc_; \
  } while (0)

int check_ranges(nbk_ctx *ctx, int n, int t0, int t1, int s0, int s1) {
  if (!ctx) return NBK_ERR_ARG;
  if (n < 0 || t0 < 0 || t1 < t0 || t1 > n || s0 < 0 || s1 < s0 || s1 > n)
    return fail(ctx, NBK_ERR_ARG, "bad ranges: n=%d targets=[%d,%d) sources=[%d,%d)", n, t0, t1,
                s0, s1);
  return NBK_OK;
}

__global__ void fill_kernel(double *p, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- sweep launcher --------------------------------------------------------------------------
template <class Op, int NT, int IPT, int MINB, int UNROLL, bool PEER = false> struct Launcher {
  static int occupancy(nbk_ctx *ctx, int *occ) {
    // the shared-memory attribute is per device: cache per device ordinal
    static int cached[64] = {0};
    int &slot = cached[ctx->device & 63];
    if (slot <= 0) {
      auto kern = sweep_kernel<Op, NT, IPT, MINB, UNROLL, PEER>;
      constexpr size_t smem = sweep_smem_bytes(naux_of<Op>::value);
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int o = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, NT, smem));
      if (o < 1) return fail(ctx, NBK_ERR_CUDA, "sweep kernel does not fit on an SM");
      slot = o;
    }
    *occ = slot;
    return NBK_OK;
  }

  // P.packed, ranges, scalars and outputs must be set; fills the tiling fields and launches.
  static int run(nbk_ctx *ctx, SweepParams P, cudaStream_t st, bool time_it) {
    constexpr int TI = NT * IPT;
    const int nt = P.t1 - P.t0, ns = P.s1 - P.s0;
    if (nt == 0) return NBK_OK;
    if (ns == 0) return fail(ctx, NBK_ERR_ARG, "empty source range");
    int occ = 0;
    TRY(occupancy(ctx, &occ));
    P.n_it = (nt + TI - 1) / TI;
    P.n_jt = (ns + TJ - 1) / TJ;
    P.units = (long long)P.n_it * P.n_jt;
    long long G = (long long)ctx->sm_count * occ;
    if (G > P.units) G = P.units;
    P.units_per_cta = (P.units + G - 1) / G;
    G = (P.units + P.units_per_cta - 1) / P.units_per_cta;  // drop empty trailing CTAs
    TRY(reserve(ctx, ctx->d_partial, (size_t)2 * G * Op::NACC * TI * sizeof(double)));
    P.partial = (double *)ctx->d_partial.p;
    if constexpr (has_tile_hook<Op>::value) {
      // the runs of a target share their time-scale bound (OpKick::tile_end) unless every
      // target's sources fit one run
      P.tbest = nullptr;
      if (P.units_per_cta < 4LL * P.n_jt && !getenv("NBK_KICK_NO_SHARED_BOUND")) {
        TRY(reserve(ctx, ctx->d_tbest, (size_t)nt * sizeof(double)));
        fill_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>((double *)ctx->d_tbest.p, (size_t)nt,
                                                                  __builtin_inf());
        CK(cudaGetLastError());
        ctx->launches++;
        P.tbest = (unsigned long long *)ctx->d_tbest.p;
      }
    }
    if (time_it) CK(cudaEventRecord(ctx->ev0, st));
    sweep_kernel<Op, NT, IPT, MINB, UNROLL, PEER>
        <<<(unsigned)G, NT, sweep_smem_bytes(naux_of<Op>::value), st>>>(P);
    CK(cudaGetLastError());
    ctx->launches++;
    if (P.units_per_cta % P.n_jt != 0) {  // some i-tile's sources are split over CTAs
      fixup_kernel<Op, TI><<<P.n_it, TI, 0, st>>>(P);
      CK(cudaGetLastError());
      ctx->launches++;
    }
    if (time_it) {
      CK(cudaEventRecord(ctx->ev1, st));
      ctx->ev_valid = true;
    }
    return NBK_OK;
  }
};

// Tile shapes.  "large": 1 CTA/SM of 256 threads x 4 targets each - two
// warps per scheduler, so consecutive FMAs of one warp issue back to back and the operand reuse
// cache works, and one source is fetched from shared memory once per 4 interactions.
// "medium"/"small" trade that for less padding when there are few targets.
enum Shape { SHAPE_AUTO = 0, SHAPE_LARGE = 1, SHAPE_MEDIUM = 2, SHAPE_SMALL = 3 };

int pick_shape(const nbk_ctx *ctx, int nt) {
  if (ctx->shape != SHAPE_AUTO) return ctx->shape;
  auto waste = [&](int ti) {
    int padded = (nt + ti - 1) / ti * ti;
    return (double)(padded - nt) / (double)padded;
  };
  if (nt >= 2048 && waste(1024) <= 0.125) return SHAPE_LARGE;
  if (nt >= 256 && waste(256) <= 0.25) return SHAPE_MEDIUM;
  return SHAPE_SMALL;
}

// Source-parallel path for a handful of targets (sweep.cuh, small_kernel).
constexpr int SMALL_NTG = 4;       // targets per launch group
constexpr int SMALL_MAX_NT = 32;   // use it up to this many targets

template <class Op> int run_small(nbk_ctx *ctx, SweepParams P, cudaStream_t st, bool time_it) {
  const int nt = P.t1 - P.t0, ns = P.s1 - P.s0;
  int ncta = (ns + 255) / 256;
  const int cap = ctx->sm_count * 4;
  if (ncta > cap) ncta = cap;
  TRY(reserve(ctx, ctx->d_partial, (size_t)nt * ncta * Op::NACC * sizeof(double)));
  P.partial = (double *)ctx->d_partial.p;
  if (time_it) CK(cudaEventRecord(ctx->ev0, st));
  for (int g0 = P.t0; g0 < P.t1; g0 += SMALL_NTG) {
    small_kernel<Op, SMALL_NTG><<<ncta, 256, 0, st>>>(P, g0, ncta);
    ctx->launches++;
  }
  CK(cudaGetLastError());
  small_fix_kernel<Op><<<(nt + 63) / 64, 64, 0, st>>>(P, ncta);
  CK(cudaGetLastError());
  ctx->launches++;
  if (time_it) {
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->ev_valid = true;
  }
  return NBK_OK;
}

template <class Op>
int run_op(nbk_ctx *ctx, const SweepParams &P, cudaStream_t st, bool time_it) {
  if (ctx->shape == SHAPE_AUTO && P.t1 - P.t0 <= SMALL_MAX_NT) return run_small<Op>(ctx, P, st, time_it);
  switch (pick_shape(ctx, P.t1 - P.t0)) {
    case SHAPE_LARGE: return Launcher<Op, 256, 4, 1, 4>::run(ctx, P, st, time_it);
    case SHAPE_MEDIUM: return Launcher<Op, 128, 2, 4, 4>::run(ctx, P, st, time_it);
    default: return Launcher<Op, 128, 1, 6, 4>::run(ctx, P, st, time_it);
  }
}

// ---- pair-once sweep (sym.cuh) ---------------------------------------------------------------
// Work items of the upper triangle for n bodies.
long long sym_items(int n) {
  const long long nb = ((long long)n + SYM_B - 1) / SYM_B;
  return nb * (nb + 1) / 2;
}
size_t sym_partial_bytes(int n) {
  const size_t nb = ((size_t)n + SYM_B - 1) / SYM_B;
  return nb * nb * SYM_B * 6 * sizeof(double);
}
// Pair-once or ordered sweep for a full n x n square on one device?  An item (1024 x 1024 pair
// evaluations, 38 FP64 instructions each) costs 38/31 of the same square of ordered pairs, the
// items come in waves of one per SM, and the ordered sweep (stream-K) balances exactly.
bool sym_pays(const nbk_ctx *ctx, int n) {
  if (ctx->sym_mode == 0 || n < 2 * SYM_B) return false;
  if (sym_partial_bytes(n) > ((size_t)24 << 30)) return false;
  if (ctx->sym_mode == 2) return true;
  const double nbf = (double)n / SYM_B;
  const long long items = sym_items(n);
  const long long waves = (items + ctx->sm_count - 1) / ctx->sm_count;
  const double t_sym = (double)waves * (38.0 / 31.0) * 1.03 + 0.02 * nbf;  // + reduce pass
  const double t_ord = nbf * nbf / ctx->sm_count;
  return t_sym < t_ord;
}

// Launches items [item0, item1) and the slot reduction.  shard/ready as in SweepParams; out6
// ([n][6], all bodies) and/or gsum/dgsum rows [t0, t1).
int sym_run(nbk_ctx *ctx, const double *const *shard, long long shard_rows, int nranks, int self,
            const unsigned long long *ready, unsigned long long epoch, int n, double eps2,
            long long item0, long long item1, double *out6, double *gsum, double *dgsum, int t0,
            int t1, cudaStream_t st, bool time_it) {
  static bool attr_set[64] = {false};
  if (!attr_set[ctx->device & 63]) {
    CK(cudaFuncSetAttribute(sym_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)SYM_SMEM));
    attr_set[ctx->device & 63] = true;
  }
  const int nb = (n + SYM_B - 1) / SYM_B;
  TRY(reserve(ctx, ctx->d_sympartial, sym_partial_bytes(n)));
  SymParams P;
  memset(&P, 0, sizeof P);
  SymReduceParams R;
  memset(&R, 0, sizeof R);
  for (int r = 0; r < nranks; ++r) P.shard[r] = R.shard[r] = shard[r];
  P.shard_rows = R.shard_rows = shard_rows;
  P.n = R.n = n;
  P.nb = R.nb = nb;
  P.nranks = nranks;
  P.self_rank = self;
  P.item0 = R.item0 = item0;
  P.item1 = R.item1 = item1;
  P.eps2 = eps2;
  P.partial = (double *)ctx->d_sympartial.p;
  P.ready = ready;
  P.epoch = epoch;
  P.wait_ns = ctx->peer_wait_ns;
  R.partial = P.partial;
  R.out6 = out6;
  R.gsum = gsum;
  R.dgsum = dgsum;
  R.t0 = t0;
  R.t1 = t1;
  if (time_it) CK(cudaEventRecord(ctx->ev0, st));
  const long long items = item1 - item0;
  if (items > 0) {
    const unsigned G = (unsigned)std::min<long long>(items, ctx->sm_count);
    sym_sweep_kernel<<<G, SYM_NT, SYM_SMEM, st>>>(P);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  if (t1 > t0) {
    sym_reduce_kernel<<<(unsigned)((t1 - t0 + 31) / 32), 192, 0, st>>>(R);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  if (time_it) {
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->ev_valid = true;
  }
  return NBK_OK;
}

// Pair-once sweeps without velocities (sym.cuh, sym_lite_kernel): acceleration only / potential
// over all n bodies of `rows`, out = [n][NACC] with the reference's sign and mass factor.
// `cost` = FP64 cost of a pair evaluation relative to an ordered square of the same size.
bool sym_lite_pays(const nbk_ctx *ctx, int n, double cost) {
  if (ctx->sym_mode == 0 || n < 2 * SYM_B) return false;
  if (sym_partial_bytes(n) > ((size_t)24 << 30)) return false;
  if (ctx->sym_mode == 2) return true;
  const double nbf = (double)n / SYM_B;
  const long long waves = (sym_items(n) + ctx->sm_count - 1) / ctx->sm_count;
  return (double)waves * cost + 0.03 * nbf < nbf * nbf / ctx->sm_count;
}

template <int OP>
int sym_lite_run(nbk_ctx *ctx, const double *rows, int n, double eps2, double *out, cudaStream_t st,
                 bool time_it) {
  constexpr int NA = SymLite<OP>::NACC;
  static bool attr_set[64] = {false};
  if (!attr_set[ctx->device & 63]) {
    CK(cudaFuncSetAttribute(sym_lite_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)SYM_LITE_SMEM));
    attr_set[ctx->device & 63] = true;
  }
  SymLiteParams P;
  memset(&P, 0, sizeof P);
  P.rows = rows;
  P.n = n;
  P.nb = (n + SYM_B - 1) / SYM_B;
  P.items = sym_items(n);
  P.eps2 = eps2;
  TRY(reserve(ctx, ctx->d_sympartial, (size_t)P.nb * P.nb * SYM_B * NA * sizeof(double)));
  P.partial = (double *)ctx->d_sympartial.p;
  P.out = out;
  if (time_it) CK(cudaEventRecord(ctx->ev0, st));
  const unsigned G = (unsigned)std::min<long long>(P.items, ctx->sm_count);
  sym_lite_kernel<OP><<<G, SYM_NT, SYM_LITE_SMEM, st>>>(P);
  CK(cudaGetLastError());
  constexpr int PER = 192 / NA;
  sym_lite_reduce_kernel<NA><<<(unsigned)((n + PER - 1) / PER), 192, 0, st>>>(P);
  CK(cudaGetLastError());
  ctx->launches += 2;
  if (time_it) {
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->ev_valid = true;
  }
  return NBK_OK;
}

// A full square over one set of bodies (targets = sources, plain store)?
bool sym_full_square(const SweepParams &P) {
  return P.t0 == P.s0 && P.t1 == P.s1 && !P.accumulate && P.out[0] != nullptr;
}

// Acceleration-only and potential sweeps: full squares with enough block pairs evaluate every
// unordered pair once (sym_lite_kernel); everything else takes the ordered sweep.
struct K1 {
  static int run(nbk_ctx *c, const SweepParams &P, cudaStream_t s, bool t) {
    if (sym_full_square(P) && sym_lite_pays(c, P.t1 - P.t0, 1.35))
      return sym_lite_run<SYM_ACC>(c, P.packed + (size_t)P.t0 * PK, P.t1 - P.t0, P.eps2, P.out[0],
                                   s, t);
    return run_op<OpGradV>(c, P, s, t);
  }
};
struct K3 {
  static int run(nbk_ctx *c, const SweepParams &P, cudaStream_t s, bool t) {
    if (sym_full_square(P) && sym_lite_pays(c, P.t1 - P.t0, 1.25))
      return sym_lite_run<SYM_POT>(c, P.packed + (size_t)P.t0 * PK, P.t1 - P.t0, P.eps2, P.out[0],
                                   s, t);
    return run_op<OpV>(c, P, s, t);
  }
};
struct K4 { static int run(nbk_ctx *c, const SweepParams &P, cudaStream_t s, bool t) { return run_op<OpKick<false>>(c, P, s, t); } };
struct K4T { static int run(nbk_ctx *c, const SweepParams &P, cudaStream_t s, bool t) { return run_op<OpKick<true>>(c, P, s, t); } };


int run_k2(nbk_ctx *ctx, const SweepParams &P, cudaStream_t st, bool time_it) {
  // A full square over one set of bodies: every unordered pair once (hermite.ml:157-166 does the
  // same), when there are enough block pairs to fill the SMs.
  if (P.t0 == P.s0 && P.t1 == P.s1 && !P.accumulate && P.out[0] && P.out[1] &&
      sym_pays(ctx, P.t1 - P.t0)) {
    const int n = P.t1 - P.t0;
    const double *rows = P.packed + (size_t)P.t0 * PK;
    const long long rows_cap = (((long long)n + SYM_B - 1) / SYM_B) * SYM_B;
    return sym_run(ctx, &rows, rows_cap, 1, 0, nullptr, 0, n, P.eps2, 0, sym_items(n), nullptr,
                   P.out[0], P.out[1], 0, n, st, time_it);
  }
  return run_op<OpGradVDGradVB<1>>(ctx, P, st, time_it);
}

// Sharded sources read from the owning GPUs (sweep_kernel<..., PEER = true>): i-parallel shapes
// only - a rank of a sharded sweep has thousands of targets.
template <class Op>
int run_op_peer(nbk_ctx *ctx, const SweepParams &P, cudaStream_t st, bool time_it) {
  switch (pick_shape(ctx, P.t1 - P.t0)) {
    case SHAPE_LARGE: return Launcher<Op, 256, 4, 1, 4, true>::run(ctx, P, st, time_it);
    case SHAPE_MEDIUM: return Launcher<Op, 128, 2, 4, 4, true>::run(ctx, P, st, time_it);
    default: return Launcher<Op, 128, 1, 6, 4, true>::run(ctx, P, st, time_it);
  }
}

// Tangent sweeps: aux tiles make the stage (1+KT) x 8 KB; one shape per KT.
template <int KT, bool JERK, bool PEER = false>
int run_tangent(nbk_ctx *ctx, const SweepParams &P, cudaStream_t st) {
  const bool time_it = (st == ctx->stream);
  if constexpr (KT == 1) {
    if (P.t1 - P.t0 >= 256)
      return Launcher<OpTangent<KT, JERK>, 128, 2, 2, 1, PEER>::run(ctx, P, st, time_it);
  }
  // KT = 3: one CTA of 256 threads per SM measured 7 % faster than two of 128 (10.09 vs 10.86 ms
  // at N = 32768; 2 targets per thread spill, deeper source unrolling gains nothing)
  if constexpr (KT == 3) {
    if (P.t1 - P.t0 >= 512)
      return Launcher<OpTangent<KT, JERK>, 256, 1, 1, 1, PEER>::run(ctx, P, st, time_it);
  }
  return Launcher<OpTangent<KT, JERK>, 128, 1, 2, 1, PEER>::run(ctx, P, st, time_it);
}

// ---- O(N) device helpers ---------------------------------------------------------------------

// Deterministic single-CTA sum of x[0..n) (fixed tree shape): Energy's final scalar.
__global__ void __launch_bounds__(1024) sum_kernel(const double *__restrict__ x, int n,
                                                   double *out) {
  __shared__ double sh[1024];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) s += x[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// ke[i] = |p_i|^2 / (2 m_i)  (energy.ml:57-58)
__global__ void kinetic_kernel(int n, const double *__restrict__ m, const double *__restrict__ p,
                               double *__restrict__ ke) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
  double n2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
  ke[i] = __ddiv_rn(n2, __dmul_rn(2.0, m[i]));
}

// Pure-DFMA throughput probe: 8 independent chains per thread.
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double seed, double *sink) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5,
         a6 = seed + 6, a7 = seed + 7;
  const double b = 1.0000001, c = 1e-9;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, b, c);
      a1 = fma(a1, b, c);
      a2 = fma(a2, b, c);
      a3 = fma(a3, b, c);
      a4 = fma(a4, b, c);
      a5 = fma(a5, b, c);
      a6 = fma(a6, b, c);
      a7 = fma(a7, b, c);
    }
  }
  double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 12345.678) sink[0] = s;
}

int upload(nbk_ctx *ctx, DevBuf &b, const double *h, size_t count) {
  TRY(reserve(ctx, b, count * sizeof(double)));
  if (count) CK(cudaMemcpyAsync(b.p, h, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return NBK_OK;
}

int download(nbk_ctx *ctx, double *h, const DevBuf &b, size_t count) {
  if (count) CK(cudaMemcpyAsync(h, b.p, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return NBK_OK;
}

int pack(nbk_ctx *ctx, int n, const double *d_m, const double *d_q, const double *d_vel,
         int is_velocity, double *d_packed, cudaStream_t st) {
  if (n == 0) return NBK_OK;
  pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, d_m, d_q, d_vel, is_velocity, d_packed);
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;
}

// Upload (m, q, vel) from the host and pack them into ctx->d_packed.
int stage_bodies(nbk_ctx *ctx, int n, const double *m, const double *q, const double *vel,
                 int is_velocity) {
  if (!m || !q) return fail(ctx, NBK_ERR_ARG, "null body array");
  TRY(upload(ctx, ctx->d_m, m, n));
  TRY(upload(ctx, ctx->d_q, q, 3 * (size_t)n));
  if (vel) TRY(upload(ctx, ctx->d_vel, vel, 3 * (size_t)n));
  TRY(reserve(ctx, ctx->d_packed, (size_t)n * PK * sizeof(double)));
  return pack(ctx, n, (const double *)ctx->d_m.p, (const double *)ctx->d_q.p,
              vel ? (const double *)ctx->d_vel.p : nullptr, is_velocity,
              (double *)ctx->d_packed.p, ctx->stream);
}

SweepParams base_params(const double *packed, double eps2, int t0, int t1, int s0, int s1) {
  SweepParams P;
  memset(&P, 0, sizeof P);
  P.packed = packed;
  P.eps2 = eps2;
  P.t0 = t0;
  P.t1 = t1;
  P.s0 = s0;
  P.s1 = s1;
  return P;
}

int sync(nbk_ctx *ctx) {
  CK(cudaStreamSynchronize(ctx->stream));
  return NBK_OK;
}

// ---- sweeps on an already packed device array, results to host ---------------------------------
int run_grad_v(nbk_ctx *ctx, const double *packed, double eps2, int t0, int t1, int s0, int s1,
               double *gsum) {
  const size_t nt = (size_t)(t1 - t0);
  if (nt == 0) return NBK_OK;
  if (!gsum) return fail(ctx, NBK_ERR_ARG, "null output");
  TRY(reserve(ctx, ctx->d_out[0], 3 * nt * sizeof(double)));
  if (s1 == s0) {
    CK(cudaMemsetAsync(ctx->d_out[0].p, 0, 3 * nt * sizeof(double), ctx->stream));
  } else {
    SweepParams P = base_params(packed, eps2, t0, t1, s0, s1);
    P.out[0] = (double *)ctx->d_out[0].p;
    TRY(K1::run(ctx, P, ctx->stream, true));
  }
  TRY(download(ctx, gsum, ctx->d_out[0], 3 * nt));
  return sync(ctx);
}

int run_grad_v_dgrad_v(nbk_ctx *ctx, const double *packed, double eps2, int t0, int t1, int s0,
                       int s1, double *gsum, double *dgsum) {
  const size_t nt = (size_t)(t1 - t0);
  if (nt == 0) return NBK_OK;
  if (!gsum || !dgsum) return fail(ctx, NBK_ERR_ARG, "null output");
  TRY(reserve(ctx, ctx->d_out[0], 3 * nt * sizeof(double)));
  TRY(reserve(ctx, ctx->d_out[1], 3 * nt * sizeof(double)));
  if (s1 == s0) {
    CK(cudaMemsetAsync(ctx->d_out[0].p, 0, 3 * nt * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_out[1].p, 0, 3 * nt * sizeof(double), ctx->stream));
  } else {
    SweepParams P = base_params(packed, eps2, t0, t1, s0, s1);
    P.out[0] = (double *)ctx->d_out[0].p;
    P.out[1] = (double *)ctx->d_out[1].p;
    TRY(run_k2(ctx, P, ctx->stream, true));
  }
  TRY(download(ctx, gsum, ctx->d_out[0], 3 * nt));
  TRY(download(ctx, dgsum, ctx->d_out[1], 3 * nt));
  return sync(ctx);
}

int run_v(nbk_ctx *ctx, const double *packed, double eps2, int t0, int t1, int s0, int s1,
          double *phi, double *total) {
  const size_t nt = (size_t)(t1 - t0);
  if (nt == 0) {
    if (total) *total = 0.0;
    return NBK_OK;
  }
  TRY(reserve(ctx, ctx->d_out[0], nt * sizeof(double)));
  if (s1 == s0) {
    CK(cudaMemsetAsync(ctx->d_out[0].p, 0, nt * sizeof(double), ctx->stream));
  } else {
    SweepParams P = base_params(packed, eps2, t0, t1, s0, s1);
    P.out[0] = (double *)ctx->d_out[0].p;
    TRY(K3::run(ctx, P, ctx->stream, true));
  }
  if (phi) TRY(download(ctx, phi, ctx->d_out[0], nt));
  if (total) {
    TRY(reserve(ctx, ctx->d_scalar, 64));
    sum_kernel<<<1, 1024, 0, ctx->stream>>>((const double *)ctx->d_out[0].p, (int)nt,
                                            (double *)ctx->d_scalar.p);
    CK(cudaGetLastError());
    ctx->launches++;
    CK(cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar.p, sizeof(double), cudaMemcpyDeviceToHost,
                       ctx->stream));
  }
  TRY(sync(ctx));
  if (total) *total = ctx->h_scalar[0];
  return NBK_OK;
}

int run_kick(nbk_ctx *ctx, const double *packed, double eta, double dt, int t0, int t1, int s0,
             int s1, double *dv, double *tmin) {
  const size_t nt = (size_t)(t1 - t0);
  if (nt == 0) return NBK_OK;
  if (!dv) return fail(ctx, NBK_ERR_ARG, "null output");
  TRY(reserve(ctx, ctx->d_out[0], 3 * nt * sizeof(double)));
  TRY(reserve(ctx, ctx->d_out[1], nt * sizeof(double)));
  if (s1 == s0) {
    CK(cudaMemsetAsync(ctx->d_out[0].p, 0, 3 * nt * sizeof(double), ctx->stream));
    if (tmin) {
      fill_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(
          (double *)ctx->d_out[1].p, nt, __builtin_inf());
      CK(cudaGetLastError());
      ctx->launches++;
    }
  } else {
    SweepParams P = base_params(packed, 0.0, t0, t1, s0, s1);
    P.dt = dt;
    P.eta = eta;
    P.out[0] = (double *)ctx->d_out[0].p;
    P.out[1] = (double *)ctx->d_out[1].p;
    if (tmin) TRY(K4T::run(ctx, P, ctx->stream, true));
    else TRY(K4::run(ctx, P, ctx->stream, true));
  }
  TRY(download(ctx, dv, ctx->d_out[0], 3 * nt));
  if (tmin) TRY(download(ctx, tmin, ctx->d_out[1], nt));
  return sync(ctx);
}

}  // namespace

// ---- single-process multi-GPU: targets sharded over the devices of a multi context ----------
// The host has all bodies.  Sharded mode (every device maps every other's memory, sources start
// on a tile): every device uploads and packs ITS OWN rows - one host->device copy per PCIe link,
// side by side - and the sweeps read the other devices' rows over NVLink themselves
// (sweep_kernel<..., PEER>); streams are ordered with events, so the kernels need no ready flags.
// One host thread per device enqueues its half (MultiPool).  Otherwise: stage on device 0 and
// broadcast the packed array (cudaMemcpyPeerAsync), serial host loop.
struct Slice {
  nbk_ctx *c;
  int a, b;  // targets [a, b) of this device
  int d;     // device index inside the multi context
};

nbk_ctx *multi_dev(nbk_ctx *ctx, int d) { return d == 0 ? ctx : ctx->peers[d - 1]; }

int multi_plan(nbk_ctx *ctx, int t0, int t1, std::vector<Slice> &out) {
  const int ndev = 1 + (int)ctx->peers.size();
  if (ctx->multi_sharded) {  // a device sweeps the targets among its own rows
    const int rows = ctx->multi_shard_rows;
    for (int d = 0; d < ndev; ++d) {
      int a = std::max(t0, d * rows), b = std::min(t1, (d + 1) * rows);
      if (b < a) b = a;
      out.push_back({multi_dev(ctx, d), a, b, d});
    }
    return NBK_OK;
  }
  const int nt = t1 - t0;
  int per = (nt + ndev - 1) / ndev;
  per = (per + 1023) / 1024 * 1024;  // whole large tiles per device
  for (int d = 0; d < ndev; ++d) {
    int a = t0 + d * per, b = t0 + (d + 1) * per;
    if (a > t1) a = t1;
    if (b > t1) b = t1;
    out.push_back({multi_dev(ctx, d), a, b, d});
  }
  return NBK_OK;
}

// Sweep parameters of one device's slice once the bodies are staged.
SweepParams multi_params(nbk_ctx *ctx, const Slice &s, double eps2, int s0, int s1) {
  SweepParams P = base_params((const double *)s.c->d_packed.p, eps2, s.a, s.b, s0, s1);
  if (ctx->multi_sharded) {
    const int ndev = 1 + (int)ctx->peers.size();
    for (int d = 0; d < ndev; ++d) P.shard[d] = (const double *)multi_dev(ctx, d)->d_packed.p;
    P.shard_rows = ctx->multi_shard_rows;
    P.self_rank = s.d;
    P.packed = P.shard[s.d] - (size_t)s.d * P.shard_rows * PK;
  }
  return P;
}

template <class Op> int multi_run(nbk_ctx *ctx, nbk_ctx *c, const SweepParams &P) {
  return ctx->multi_sharded ? run_op_peer<Op>(c, P, c->stream, true)
                            : run_op<Op>(c, P, c->stream, true);
}

// Runs fn(d) for every device of the multi context, device d on its own host thread (d = 0 on
// the caller's), and returns when all are done.  Without a pool: a serial loop.
void multi_for(nbk_ctx *ctx, const std::function<void(int)> &fn) {
  const int ndev = 1 + (int)ctx->peers.size();
  MultiPool *pool = ctx->pool;
  if (!pool) {
    for (int d = 0; d < ndev; ++d) fn(d);
    return;
  }
  {
    std::lock_guard<std::mutex> lk(pool->mu);
    pool->job = fn;
    pool->pending = ndev - 1;
    pool->generation++;
  }
  pool->cv_go.notify_all();
  fn(0);
  std::unique_lock<std::mutex> lk(pool->mu);
  pool->cv_done.wait(lk, [&] { return pool->pending == 0; });
}

void multi_worker(MultiPool *pool, int d) {
  unsigned long long seen = 0;
  for (;;) {
    std::function<void(int)> job;
    {
      std::unique_lock<std::mutex> lk(pool->mu);
      pool->cv_go.wait(lk, [&] { return pool->stop || pool->generation != seen; });
      if (pool->stop) return;
      seen = pool->generation;
      job = pool->job;
    }
    job(d);
    {
      std::lock_guard<std::mutex> lk(pool->mu);
      if (--pool->pending == 0) pool->cv_done.notify_all();
    }
  }
}

// One sharded sweep of a multi context from host buffers.  `launch(c, slice, P)` enqueues the
// sweep of the device's slice on c->stream, `collect(c, slice)` its read-backs (both called with
// the device current, possibly on a worker thread; they must touch only c).  The serial paths
// launch everywhere before collecting anywhere - a read-back into pageable memory blocks the
// host.  Returns after every device has finished.
using MultiLaunch = std::function<int(nbk_ctx *, const Slice &, SweepParams)>;
using MultiCollect = std::function<int(nbk_ctx *, const Slice &)>;
// optional: stage more per-row data of rows [a, b) on device c (sharded mode only; `rows` = rows
// per device), e.g. the tangent directions
using MultiStage = std::function<int(nbk_ctx *, int a, int b, int rows)>;

// Sharded mode needs every device to map every other's memory, sources starting on a tile, and
// at least one tile per device.
bool multi_can_shard(const nbk_ctx *ctx, int n, int s0) {
  const int ndev = 1 + (int)ctx->peers.size();
  return ctx->p2p_all && s0 % TJ == 0 && n >= ndev * TJ;
}

int multi_sweep(nbk_ctx *ctx, int n, const double *m, const double *q, const double *vel,
                int is_velocity, double eps2, int t0, int t1, int s0, int s1,
                const MultiLaunch &launch, const MultiCollect &collect,
                const MultiStage &stage = nullptr) {
  const int ndev = 1 + (int)ctx->peers.size();
  if (!m || !q) return fail(ctx, NBK_ERR_ARG, "null body array");
  ctx->multi_sharded = multi_can_shard(ctx, n, s0);
  if (!ctx->multi_sharded) {
    if (stage) return fail(ctx, NBK_ERR_STATE, "multi_sweep: extra row data needs the sharded mode");
    // fallback: stage on device 0, broadcast, serial host loop
    CK(cudaSetDevice(ctx->device));
    TRY(stage_bodies(ctx, n, m, q, vel, is_velocity));
    CK(cudaEventRecord(ctx->ev_ready, ctx->stream));
    const size_t bytes = (size_t)n * PK * sizeof(double);
    for (nbk_ctx *peer : ctx->peers) {
      CK(cudaSetDevice(peer->device));
      TRY(reserve(peer, peer->d_packed, bytes));
      CK(cudaStreamWaitEvent(peer->stream, ctx->ev_ready, 0));
      CK(cudaMemcpyPeerAsync(peer->d_packed.p, peer->device, ctx->d_packed.p, ctx->device, bytes,
                             peer->stream));
    }
    std::vector<Slice> sl;
    TRY(multi_plan(ctx, t0, t1, sl));
    for (Slice &s : sl) {  // launch everywhere first
      if (s.b == s.a) continue;
      if (cudaSetDevice(s.c->device) != cudaSuccess) return fail(ctx, NBK_ERR_CUDA, "cudaSetDevice");
      int rc = launch(s.c, s, multi_params(ctx, s, eps2, s0, s1));
      if (rc != NBK_OK) return fail(ctx, rc, "%s", nbk_last_error(s.c));
    }
    for (Slice &s : sl) {
      cudaSetDevice(s.c->device);
      int rc = s.b > s.a ? collect(s.c, s) : NBK_OK;
      if (rc == NBK_OK) rc = sync(s.c);
      if (rc != NBK_OK) return fail(ctx, rc, "%s", nbk_last_error(s.c));
    }
    CK(cudaSetDevice(ctx->device));
    return NBK_OK;
  }
  int rows = (n + ndev - 1) / ndev;
  rows = (rows + 1023) / 1024 * 1024;  // whole large target tiles (and source tiles) per device
  ctx->multi_shard_rows = rows;
  std::vector<Slice> sl;
  TRY(multi_plan(ctx, t0, t1, sl));
  std::vector<int> rcs(ndev, NBK_OK);
  if (ctx->pool) ctx->pool->nparticipants = ndev;
  // phase 1: device d uploads and packs its own rows, then records "rows packed"
  auto pack_rows = [&](int d) {
    nbk_ctx *c = multi_dev(ctx, d);
    int rc = NBK_OK;
    if (cudaSetDevice(c->device) != cudaSuccess) rc = fail(c, NBK_ERR_CUDA, "cudaSetDevice failed");
    const int a = std::min(n, d * rows), b = std::min(n, (d + 1) * rows), cnt = b - a;
    if (rc == NBK_OK) rc = reserve(c, c->d_packed, (size_t)rows * PK * sizeof(double));
    if (rc == NBK_OK && cnt > 0) {
      rc = upload(c, c->d_m, m + a, cnt);
      if (rc == NBK_OK) rc = upload(c, c->d_q, q + 3 * (size_t)a, 3 * (size_t)cnt);
      if (rc == NBK_OK && vel) rc = upload(c, c->d_vel, vel + 3 * (size_t)a, 3 * (size_t)cnt);
      if (rc == NBK_OK)
        rc = pack(c, cnt, (const double *)c->d_m.p, (const double *)c->d_q.p,
                  vel ? (const double *)c->d_vel.p : nullptr, is_velocity, (double *)c->d_packed.p,
                  c->stream);
    }
    if (rc == NBK_OK && stage) rc = stage(c, a, b, rows);
    if (rc == NBK_OK && cudaEventRecord(c->ev_ready, c->stream) != cudaSuccess)
      rc = fail(c, NBK_ERR_CUDA, "cudaEventRecord failed");
    rcs[d] = rc;
  };
  auto all_packed = [&]() {
    for (int e = 0; e < ndev; ++e)
      if (rcs[e] != NBK_OK) return false;
    return true;
  };
  // phase 2: nobody sweeps before every shard is packed
  auto sweep_slice = [&](int d) {
    nbk_ctx *c = multi_dev(ctx, d);
    int rc = NBK_OK;
    if (cudaSetDevice(c->device) != cudaSuccess) rc = fail(c, NBK_ERR_CUDA, "cudaSetDevice failed");
    for (int e = 0; e < ndev && rc == NBK_OK; ++e)
      if (e != d && cudaStreamWaitEvent(c->stream, multi_dev(ctx, e)->ev_ready, 0) != cudaSuccess)
        rc = fail(c, NBK_ERR_CUDA, "cudaStreamWaitEvent failed");
    const Slice &s = sl[d];
    if (rc == NBK_OK && s.b > s.a) rc = launch(c, s, multi_params(ctx, s, eps2, s0, s1));
    rcs[d] = rc;
  };
  // phase 3: the slice goes home as soon as it is done
  auto collect_slice = [&](int d, bool swept) {
    nbk_ctx *c = multi_dev(ctx, d);
    cudaSetDevice(c->device);
    const Slice &s = sl[d];
    if (swept && rcs[d] == NBK_OK && s.b > s.a) rcs[d] = collect(c, s);
    const int rc2 = sync(c);
    if (rcs[d] == NBK_OK) rcs[d] = rc2;
  };
  if (ctx->pool) {
    // one host thread per device; the pool barrier separates phase 1 from phase 2
    multi_for(ctx, [&](int d) {
      pack_rows(d);
      ctx->pool->barrier();  // every device's "rows packed" event is recorded (or failed) past here
      const bool ok = all_packed();
      if (ok) sweep_slice(d);
      collect_slice(d, ok);
    });
  } else {
    // serial host loop: one pass per phase stands in for the barrier, and every device's sweep is
    // in flight before the first blocking download
    for (int d = 0; d < ndev; ++d) pack_rows(d);
    const bool ok = all_packed();
    for (int d = 0; d < ndev && ok; ++d) sweep_slice(d);
    for (int d = 0; d < ndev; ++d) collect_slice(d, ok);
  }
  cudaSetDevice(ctx->device);
  for (int d = 0; d < ndev; ++d)
    if (rcs[d] != NBK_OK) {
      nbk_ctx *c = multi_dev(ctx, d);
      return c == ctx ? rcs[d] : fail(ctx, rcs[d], "device %d: %s", c->device, nbk_last_error(c));
    }
  return NBK_OK;
}

// The full acc+jerk square of a multi context with every unordered pair evaluated once on the
// whole box (sym.cuh): device d uploads and packs its own rows, runs its contiguous range of
// block pairs (rows of both blocks read from their owners over NVLink) and reduces its slots to
// sums for all bodies; after a second barrier it adds the devices' sums of ITS bodies in device
// order (peer reads) and sends the slice home.  Streams are ordered with events.
int multi_sym_acc_jerk(nbk_ctx *ctx, int n, const double *m, const double *q, const double *p,
                       double eps2, double *gsum, double *dgsum) {
  const int ndev = 1 + (int)ctx->peers.size();
  int rows = (n + ndev - 1) / ndev;
  rows = (rows + SYM_B - 1) / SYM_B * SYM_B;
  ctx->multi_sharded = true;
  ctx->multi_shard_rows = rows;
  const long long items = sym_items(n);
  std::vector<int> rcs(ndev, NBK_OK);
  if (ctx->pool) ctx->pool->nparticipants = ndev;
  auto ok_all = [&]() {
    for (int e = 0; e < ndev; ++e)
      if (rcs[e] != NBK_OK) return false;
    return true;
  };
  auto pack_rows = [&](int d) {
    nbk_ctx *c = multi_dev(ctx, d);
    int rc = NBK_OK;
    if (cudaSetDevice(c->device) != cudaSuccess) rc = fail(c, NBK_ERR_CUDA, "cudaSetDevice failed");
    const int a = std::min(n, d * rows), b = std::min(n, (d + 1) * rows), cnt = b - a;
    if (rc == NBK_OK) rc = reserve(c, c->d_packed, (size_t)rows * PK * sizeof(double));
    if (rc == NBK_OK) rc = reserve(c, c->d_out[2], (size_t)n * 6 * sizeof(double));
    if (rc == NBK_OK) rc = reserve(c, c->d_sympartial, sym_partial_bytes(n));
    if (rc == NBK_OK && cnt > 0) {
      rc = upload(c, c->d_m, m + a, cnt);
      if (rc == NBK_OK) rc = upload(c, c->d_q, q + 3 * (size_t)a, 3 * (size_t)cnt);
      if (rc == NBK_OK) rc = upload(c, c->d_vel, p + 3 * (size_t)a, 3 * (size_t)cnt);
      if (rc == NBK_OK)
        rc = pack(c, cnt, (const double *)c->d_m.p, (const double *)c->d_q.p,
                  (const double *)c->d_vel.p, 0, (double *)c->d_packed.p, c->stream);
    }
    if (rc == NBK_OK && cudaEventRecord(c->ev_ready, c->stream) != cudaSuccess)
      rc = fail(c, NBK_ERR_CUDA, "cudaEventRecord failed");
    rcs[d] = rc;
  };
  auto sweep_items = [&](int d) {
    nbk_ctx *c = multi_dev(ctx, d);
    int rc = NBK_OK;
    if (cudaSetDevice(c->device) != cudaSuccess) rc = fail(c, NBK_ERR_CUDA, "cudaSetDevice failed");
    for (int e = 0; e < ndev && rc == NBK_OK; ++e)
      if (e != d && cudaStreamWaitEvent(c->stream, multi_dev(ctx, e)->ev_ready, 0) != cudaSuccess)
        rc = fail(c, NBK_ERR_CUDA, "cudaStreamWaitEvent failed");
    const double *shard[MAX_PEERS];
    for (int e = 0; e < ndev; ++e) shard[e] = (const double *)multi_dev(ctx, e)->d_packed.p;
    if (rc == NBK_OK)
      rc = sym_run(c, shard, rows, ndev, d, nullptr, 0, n, eps2, items * d / ndev,
                   items * (d + 1) / ndev, (double *)c->d_out[2].p, nullptr, nullptr, 0, n,
                   c->stream, true);
    if (rc == NBK_OK && cudaEventRecord(c->ev_sum, c->stream) != cudaSuccess)
      rc = fail(c, NBK_ERR_CUDA, "cudaEventRecord failed");
    rcs[d] = rc;
  };
  auto gather_slice = [&](int d, bool swept) {
    nbk_ctx *c = multi_dev(ctx, d);
    cudaSetDevice(c->device);
    int rc = rcs[d];
    const int a = std::min(n, d * rows), b = std::min(n, (d + 1) * rows), cnt = b - a;
    if (swept && rc == NBK_OK && cnt > 0) {
      for (int e = 0; e < ndev && rc == NBK_OK; ++e)
        if (e != d && cudaStreamWaitEvent(c->stream, multi_dev(ctx, e)->ev_sum, 0) != cudaSuccess)
          rc = fail(c, NBK_ERR_CUDA, "cudaStreamWaitEvent failed");
      if (rc == NBK_OK) rc = reserve(c, c->d_out[0], (size_t)cnt * 3 * sizeof(double));
      if (rc == NBK_OK) rc = reserve(c, c->d_out[1], (size_t)cnt * 3 * sizeof(double));
      if (rc == NBK_OK) {
        SymGatherParams G;
        memset(&G, 0, sizeof G);
        for (int e = 0; e < ndev; ++e) G.sum6[e] = (const double *)multi_dev(ctx, e)->d_out[2].p;
        G.ndev = ndev;
        G.self = d;
        G.a = a;
        G.b = b;
        G.g = (double *)c->d_out[0].p;
        G.dg = (double *)c->d_out[1].p;
        sym_gather_kernel<<<(cnt * 6 + 255) / 256, 256, 0, c->stream>>>(G);
        if (cudaGetLastError() != cudaSuccess) rc = fail(c, NBK_ERR_CUDA, "sym_gather_kernel launch failed");
        c->launches++;
      }
      if (rc == NBK_OK) rc = download(c, gsum + 3 * (size_t)a, c->d_out[0], 3 * (size_t)cnt);
      if (rc == NBK_OK) rc = download(c, dgsum + 3 * (size_t)a, c->d_out[1], 3 * (size_t)cnt);
    }
    const int rc2 = sync(c);
    rcs[d] = rc != NBK_OK ? rc : rc2;
  };
  if (ctx->pool) {
    multi_for(ctx, [&](int d) {
      pack_rows(d);
      ctx->pool->barrier();
      const bool ok = ok_all();
      if (ok) sweep_items(d);
      ctx->pool->barrier();  // every device's "sums complete" event is recorded (or failed)
      gather_slice(d, ok && ok_all());
    });
  } else {
    for (int d = 0; d < ndev; ++d) pack_rows(d);
    const bool ok = ok_all();
    for (int d = 0; d < ndev && ok; ++d) sweep_items(d);
    const bool ok2 = ok && ok_all();
    // every device's gather is enqueued before the first blocking read-back completes
    for (int d = 0; d < ndev; ++d) gather_slice(d, ok2);
  }
  cudaSetDevice(ctx->device);
  for (int d = 0; d < ndev; ++d)
    if (rcs[d] != NBK_OK) {
      nbk_ctx *c = multi_dev(ctx, d);
      return c == ctx ? rcs[d] : fail(ctx, rcs[d], "device %d: %s", c->device, nbk_last_error(c));
    }
  return NBK_OK;
}

// =================================================================================================
extern "C" {

const char *nbk_version(void) { return "libnbk 0.1 (sm_100a)"; }

const char *nbk_last_error(const nbk_ctx *ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int nbk_create(nbk_ctx **out, int device) {
  nbk_ctx *ctx = nullptr;  // CK() reports into g_create_error while ctx == NULL
  if (!out) return fail(nullptr, NBK_ERR_ARG, "nbk_create: null out pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, NBK_ERR_CUDA,
                "nbk_create: no CUDA device (%s); libnbk has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count)
    return fail(nullptr, NBK_ERR_ARG, "nbk_create: device %d out of range [0,%d)", device, count);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, NBK_ERR_CUDA,
                "nbk_create: device %d is sm_%d%d; libnbk is built for sm_100a (B200) only",
                device, prop.major, prop.minor);
  nbk_ctx *c = new (std::nothrow) nbk_ctx();
  if (!c) return fail(nullptr, NBK_ERR_NOMEM, "nbk_create: out of host memory");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (const char *e = getenv("NBK_HERMITE_THREADS")) c->hermite_threads = atoi(e);
  if (const char *e = getenv("NBK_HERMITE_SPEC")) c->hermite_spec = atoi(e) != 0;
  if (const char *e = getenv("NBK_SYM")) c->sym_mode = atoi(e);
  if (const char *e = getenv("NBK_PEER_TIMEOUT_S")) {
    const double s = atof(e);
    c->peer_wait_ns = s > 0.0 ? (unsigned long long)(s * 1e9) : 0ULL;
  }
  cudaError_t e1 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  cudaError_t e2 = cudaEventCreate(&c->ev0);
  cudaError_t e3 = cudaEventCreate(&c->ev1);
  cudaError_t e4 = cudaMallocHost((void **)&c->h_scalar, 64);
  cudaError_t e5 = cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming);
  if (e5 == cudaSuccess) e5 = cudaEventCreateWithFlags(&c->ev_sum, cudaEventDisableTiming);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess ||
      e5 != cudaSuccess) {
    nbk_destroy(c);
    return fail(nullptr, NBK_ERR_CUDA, "nbk_create: stream/event/pinned allocation failed");
  }
  *out = c;
  return NBK_OK;
}

int nbk_create_multi(nbk_ctx **out, int ndev, const int *devices) {
  nbk_ctx *ctx = nullptr;
  if (!out || ndev < 1) return fail(nullptr, NBK_ERR_ARG, "nbk_create_multi: bad arguments");
  *out = nullptr;
  nbk_ctx *root = nullptr;
  TRY(nbk_create(&root, devices ? devices[0] : 0));
  for (int d = 1; d < ndev; ++d) {
    nbk_ctx *peer = nullptr;
    int rc = nbk_create(&peer, devices ? devices[d] : d);
    if (rc != NBK_OK) {
      nbk_destroy(root);
      return rc;
    }
    root->peers.push_back(peer);
  }
  // direct NVLink access between every pair of devices (ignore "already enabled")
  root->p2p_all = ndev > 1 && getenv("NBK_NO_PEER_READS") == nullptr;
  for (int a = 0; a < ndev; ++a)
    for (int b = 0; b < ndev; ++b) {
      if (a == b) continue;
      nbk_ctx *ca = multi_dev(root, a), *cb = multi_dev(root, b);
      if (ca->device == cb->device) continue;  // same GPU listed twice (tests): trivially mapped
      int can = 0;
      cudaDeviceCanAccessPeer(&can, ca->device, cb->device);
      if (can) {
        cudaSetDevice(ca->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(cb->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
        cudaGetLastError();
      }
      if (!can) root->p2p_all = false;
    }
  CK(cudaSetDevice(root->device));
  // one host thread per extra device (NBK_MULTI_THREADS=0: serial host loop)
  const char *mt = getenv("NBK_MULTI_THREADS");
  if (ndev > 1 && !(mt && atoi(mt) == 0)) {
    root->pool = new (std::nothrow) MultiPool();
    if (root->pool) {
      root->pool->nparticipants = ndev;
      for (int d = 1; d < ndev; ++d) root->pool->threads.emplace_back(multi_worker, root->pool, d);
    }
  }
  *out = root;
  return NBK_OK;
}

int nbk_device_count(const nbk_ctx *ctx) { return ctx ? 1 + (int)ctx->peers.size() : 0; }

void nbk_destroy(nbk_ctx *c) {
  if (!c) return;
  if (c->pool) {
    {
      std::lock_guard<std::mutex> lk(c->pool->mu);
      c->pool->stop = true;
    }
    c->pool->cv_go.notify_all();
    for (std::thread &t : c->pool->threads) t.join();
    delete c->pool;
    c->pool = nullptr;
  }
  for (nbk_ctx *peer : c->peers) nbk_destroy(peer);
  c->peers.clear();
  cudaSetDevice(c->device);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->ev_sum) cudaEventDestroy(c->ev_sum);
  if (c->stream) cudaStreamSynchronize(c->stream);
  DevBuf *bufs[] = {&c->d_t,     &c->d_f,      &c->d_df,     &c->d_aux_in2,
                    &c->d_m,     &c->d_q,      &c->d_vel,    &c->d_aux_in, &c->d_packed,
                    &c->d_aux,   &c->d_partial, &c->d_tbest, &c->d_out[0], &c->d_out[1], &c->d_out[2],
                    &c->d_out[3], &c->d_scalar, &c->d_res_m,  &c->d_res_packed,
                    &c->d_hstate, &c->d_hpartial, &c->d_hticket, &c->d_hout, &c->d_hgrid,
                    &c->d_adv,    &c->d_adv2,   &c->d_oa_m,   &c->d_oa_q,  &c->d_oa_p,
                    &c->d_oa_qs,  &c->d_oa_g,   &c->d_oa_phi, &c->d_oa_el, &c->d_lf_packed,
                    &c->d_lf_packed2, &c->d_lf_dtmax, &c->d_lf_dtmax2, &c->d_adt, &c->d_lft, &c->d_sympartial};
  for (DevBuf *b : bufs)
    if (b->p) cudaFree(b->p);
  if (c->h_scalar) cudaFreeHost(c->h_scalar);
  if (c->h_hres) cudaFreeHost(c->h_hres);
  if (c->h_lf_stage) cudaFreeHost(c->h_lf_stage);
  if (c->ev_lf) cudaEventDestroy(c->ev_lf);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int nbk_sm_count(const nbk_ctx *ctx) { return ctx ? ctx->sm_count : 0; }
int64_t nbk_launch_count(const nbk_ctx *ctx) {
  if (!ctx) return 0;
  int64_t tot = ctx->launches;
  for (const nbk_ctx *peer : ctx->peers) tot += peer->launches;
  return tot;
}

int nbk_last_kernel_ms(nbk_ctx *ctx, float *ms) {
  if (!ctx || !ms) return NBK_ERR_ARG;
  if (!ctx->ev_valid) return fail(ctx, NBK_ERR_STATE, "no sweep has been timed yet");
  CK(cudaEventSynchronize(ctx->ev1));
  CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return NBK_OK;
}

// ---- host-buffer sweeps ----------------------------------------------------------------------
int nbk_grad_v_sum(nbk_ctx *ctx, int n, const double *m, const double *q, double eps2, int t0,
                   int t1, int s0, int s1, double *gsum) {
  TRY(check_ranges(ctx, n, t0, t1, s0, s1));
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return NBK_OK;
  if (!ctx->peers.empty() && t1 > t0 && s1 > s0) {
    if (!gsum) return fail(ctx, NBK_ERR_ARG, "null output");
    return multi_sweep(ctx, n, m, q, nullptr, 1, eps2, t0, t1, s0, s1,
                       [=](nbk_ctx *c, const Slice &sl, SweepParams P) -> int {
                         const size_t cnt = 3 * (size_t)(sl.b - sl.a);
                         int rc = reserve(c, c->d_out[0], cnt * sizeof(double));
                         if (rc != NBK_OK) return rc;
                         P.out[0] = (double *)c->d_out[0].p;
                         return multi_run<OpGradV>(ctx, c, P);
                       },
                       [=](nbk_ctx *c, const Slice &sl) -> int {
                         return download(c, gsum + 3 * (size_t)(sl.a - t0), c->d_out[0],
                                         3 * (size_t)(sl.b - sl.a));
                       });
  }
  TRY(stage_bodies(ctx, n, m, q, nullptr, 1));
  return run_grad_v(ctx, (const double *)ctx->d_packed.p, eps2, t0, t1, s0, s1, gsum);
}

int nbk_grad_v_dgrad_v_sum(nbk_ctx *ctx, int n, const double *m, const double *q,
                           const double *p, double eps2, int t0, int t1, int s0, int s1,
                           double *gsum, double *dgsum) {
  TRY(check_ranges(ctx, n, t0, t1, s0, s1));
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return NBK_OK;
  if (!p) return fail(ctx, NBK_ERR_ARG, "null momentum array");
  if (!ctx->peers.empty() && t1 > t0 && s1 > s0) {
    if (!gsum || !dgsum) return fail(ctx, NBK_ERR_ARG, "null output");
    // the full square on a box whose devices map each other's memory: every pair once
    if (t0 == 0 && s0 == 0 && t1 == n && s1 == n && ctx->p2p_all && ctx->sym_mode != 0 &&
        n >= (int)(1 + ctx->peers.size()) * SYM_B &&
        (ctx->sym_mode == 2 || sym_items(n) >= 4LL * ctx->sm_count * (1 + (long long)ctx->peers.size())) &&
        sym_partial_bytes(n) <= ((size_t)24 << 30))
      return multi_sym_acc_jerk(ctx, n, m, q, p, eps2, gsum, dgsum);
    return multi_sweep(ctx, n, m, q, p, 0, eps2, t0, t1, s0, s1,
                       [=](nbk_ctx *c, const Slice &sl, SweepParams P) -> int {
                         const size_t cnt = 3 * (size_t)(sl.b - sl.a);
                         int rc = reserve(c, c->d_out[0], cnt * sizeof(double));
                         if (rc == NBK_OK) rc = reserve(c, c->d_out[1], cnt * sizeof(double));
                         if (rc != NBK_OK) return rc;
                         P.out[0] = (double *)c->d_out[0].p;
                         P.out[1] = (double *)c->d_out[1].p;
                         return multi_run<OpGradVDGradVB<1>>(ctx, c, P);
                       },
                       [=](nbk_ctx *c, const Slice &sl) -> int {
                         const size_t cnt = 3 * (size_t)(sl.b - sl.a);
                         int rc = download(c, gsum + 3 * (size_t)(sl.a - t0), c->d_out[0], cnt);
                         if (rc != NBK_OK) return rc;
                         return download(c, dgsum + 3 * (size_t)(sl.a - t0), c->d_out[1], cnt);
                       });
  }
  TRY(stage_bodies(ctx, n, m, q, p, 0));
  return run_grad_v_dgrad_v(ctx, (const double *)ctx->d_packed.p, eps2, t0, t1, s0, s1, gsum,
                            dgsum);
}

int nbk_v_sum(nbk_ctx *ctx, int n, const double *m, const double *q, double eps2, int t0, int t1,
              int s0, int s1, double *phi, double *total) {
  TRY(check_ranges(ctx, n, t0, t1, s0, s1));
  CK(cudaSetDevice(ctx->device));
  if (n == 0) {
    if (total) *total = 0.0;
    return NBK_OK;
  }
  if (!ctx->peers.empty() && t1 > t0 && s1 > s0) {
    const int ndev = 1 + (int)ctx->peers.size();
    std::vector<char> has(ndev, 0);
    char *hasp = has.data();
    TRY(multi_sweep(ctx, n, m, q, nullptr, 1, eps2, t0, t1, s0, s1,
                    [=](nbk_ctx *c, const Slice &sl, SweepParams P) -> int {
                      const size_t cnt = (size_t)(sl.b - sl.a);
                      int rc = reserve(c, c->d_out[0], cnt * sizeof(double));
                      if (rc == NBK_OK) rc = reserve(c, c->d_scalar, 64);
                      if (rc != NBK_OK) return rc;
                      P.out[0] = (double *)c->d_out[0].p;
                      rc = multi_run<OpV>(ctx, c, P);
                      if (rc != NBK_OK) return rc;
                      if (total) {
                        sum_kernel<<<1, 1024, 0, c->stream>>>((const double *)c->d_out[0].p,
                                                              (int)cnt, (double *)c->d_scalar.p);
                        c->launches++;
                        if (cudaMemcpyAsync(c->h_scalar, c->d_scalar.p, sizeof(double),
                                            cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
                          return fail(c, NBK_ERR_CUDA, "cudaMemcpyAsync failed");
                        hasp[sl.d] = 1;
                      }
                      return NBK_OK;
                    },
                    [=](nbk_ctx *c, const Slice &sl) -> int {
                      if (phi)
                        return download(c, phi + (size_t)(sl.a - t0), c->d_out[0],
                                        (size_t)(sl.b - sl.a));
                      return NBK_OK;
                    }));
    if (total) {  // per-device partial sums, combined in device order (fixed summation order)
      double tot = 0.0;
      for (int d = 0; d < ndev; ++d)
        if (has[d]) tot += multi_dev(ctx, d)->h_scalar[0];
      *total = tot;
    }
    return NBK_OK;
  }
  TRY(stage_bodies(ctx, n, m, q, nullptr, 1));
  return run_v(ctx, (const double *)ctx->d_packed.p, eps2, t0, t1, s0, s1, phi, total);
}

int nbk_energy(nbk_ctx *ctx, int n, const double *m, const double *q, const double *p,
               double eps2, double *ke, double *pe) {
  if (!ctx) return NBK_ERR_ARG;
  if (n < 0) return fail(ctx, NBK_ERR_ARG, "n < 0");
  CK(cudaSetDevice(ctx->device));
  if (ke) *ke = 0.0;
  if (pe) *pe = 0.0;
  if (n == 0) return NBK_OK;
  if (ke && !p) return fail(ctx, NBK_ERR_ARG, "kinetic energy needs momenta");
  if (pe) {
    double tot = 0.0;
    TRY(nbk_v_sum(ctx, n, m, q, eps2, 0, n, 0, n, nullptr, &tot));
    *pe = 0.5 * tot;
  }
  if (ke) {
    TRY(upload(ctx, ctx->d_m, m, n));
    TRY(upload(ctx, ctx->d_vel, p, 3 * (size_t)n));
    TRY(reserve(ctx, ctx->d_out[0], (size_t)n * sizeof(double)));
    TRY(reserve(ctx, ctx->d_scalar, 64));
    kinetic_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(
        n, (const double *)ctx->d_m.p, (const double *)ctx->d_vel.p, (double *)ctx->d_out[0].p);
    CK(cudaGetLastError());
    sum_kernel<<<1, 1024, 0, ctx->stream>>>((const double *)ctx->d_out[0].p, n,
                                            (double *)ctx->d_scalar.p);
    CK(cudaGetLastError());
    ctx->launches += 2;
    CK(cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar.p, sizeof(double), cudaMemcpyDeviceToHost,
                       ctx->stream));
    TRY(sync(ctx));
    *ke = ctx->h_scalar[0];
  }
  return NBK_OK;
}

int nbk_kick_sum(nbk_ctx *ctx, int n, const double *m, const double *q, const double *v,
                 double eta, double dt, int t0, int t1, int s0, int s1, double *dv,
                 double *tmin) {
  TRY(check_ranges(ctx, n, t0, t1, s0, s1));
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return NBK_OK;
  if (tmin && !v) return fail(ctx, NBK_ERR_ARG, "time-scale needs velocities");
  if (!ctx->peers.empty() && t1 > t0 && s1 > s0) {  // targets sharded over the devices
    if (!dv) return fail(ctx, NBK_ERR_ARG, "null output");
    return multi_sweep(ctx, n, m, q, v, 1, 0.0, t0, t1, s0, s1,
                       [=](nbk_ctx *c, const Slice &sl, SweepParams P) -> int {
                         const size_t cnt = (size_t)(sl.b - sl.a);
                         int rc = reserve(c, c->d_out[0], 3 * cnt * sizeof(double));
                         if (rc == NBK_OK) rc = reserve(c, c->d_out[1], cnt * sizeof(double));
                         if (rc != NBK_OK) return rc;
                         P.dt = dt;
                         P.eta = eta;
                         P.out[0] = (double *)c->d_out[0].p;
                         P.out[1] = (double *)c->d_out[1].p;
                         return tmin ? multi_run<OpKick<true>>(ctx, c, P)
                                     : multi_run<OpKick<false>>(ctx, c, P);
                       },
                       [=](nbk_ctx *c, const Slice &sl) -> int {
                         const size_t cnt = (size_t)(sl.b - sl.a);
                         int rc = download(c, dv + 3 * (size_t)(sl.a - t0), c->d_out[0], 3 * cnt);
                         if (rc == NBK_OK && tmin)
                           rc = download(c, tmin + (size_t)(sl.a - t0), c->d_out[1], cnt);
                         return rc;
                       });
  }
  TRY(stage_bodies(ctx, n, m, q, v, 1));
  return run_kick(ctx, (const double *)ctx->d_packed.p, eta, dt, t0, t1, s0, s1, dv, tmin);
}

int nbk_binary_scan(nbk_ctx *ctx, int n, const double *m, const double *q, const double *p,
                    double eps2, double threshold, int t0, int t1, int s0, int s1, double *emin,
                    int *partner, int *count) {
  TRY(check_ranges(ctx, n, t0, t1, s0, s1));
  CK(cudaSetDevice(ctx->device));
  const int nt = t1 - t0;
  if (n == 0 || nt == 0) return NBK_OK;
  if (!p) return fail(ctx, NBK_ERR_ARG, "null momentum array");
  TRY(stage_bodies(ctx, n, m, q, p, 0));
  index_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((double *)ctx->d_packed.p, n);
  CK(cudaGetLastError());
  ctx->launches++;
  for (int k = 0; k < 3; ++k) TRY(reserve(ctx, ctx->d_out[k], (size_t)nt * sizeof(double)));
  std::string hbuf((size_t)nt * 3 * sizeof(double), '\0');
  double *h = reinterpret_cast<double *>(&hbuf[0]);
  if (s1 == s0) {
    for (int i = 0; i < nt; ++i) {
      h[i] = __builtin_inf();
      h[nt + i] = -1.0;
      h[2 * nt + i] = 0.0;
    }
  } else {
    SweepParams P = base_params((const double *)ctx->d_packed.p, eps2, t0, t1, s0, s1);
    P.dt = threshold;
    for (int k = 0; k < 3; ++k) P.out[k] = (double *)ctx->d_out[k].p;
    TRY(run_op<OpBinary>(ctx, P, ctx->stream, true));
    for (int k = 0; k < 3; ++k)
      CK(cudaMemcpyAsync(h + (size_t)k * nt, ctx->d_out[k].p, (size_t)nt * sizeof(double),
                         cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync(ctx));
  }
  for (int i = 0; i < nt; ++i) {
    if (emin) emin[i] = h[i];
    if (partner) partner[i] = (int)h[nt + i];
    if (count) count[i] = (int)h[2 * nt + i];
  }
  return NBK_OK;
}

int nbk_binary_list(nbk_ctx *ctx, int n, const double *m, const double *q, const double *p,
                    double eps2, double threshold, int max_pairs, int *pi, int *pj, double *e,
                    int *npairs) {
  if (!ctx) return NBK_ERR_ARG;
  if (n < 0 || max_pairs < 0 || !npairs || (max_pairs > 0 && (!pi || !pj || !e)))
    return fail(ctx, NBK_ERR_ARG, "nbk_binary_list: bad arguments");
  *npairs = 0;
  if (n < 2) return NBK_OK;
  std::string cbuf((size_t)n * sizeof(int), '\0');
  int *cnt = reinterpret_cast<int *>(&cbuf[0]);
  TRY(nbk_binary_scan(ctx, n, m, q, p, eps2, threshold, 0, n, 0, n, nullptr, nullptr, cnt));
  std::string fbuf((size_t)n * sizeof(int), '\0');
  int *flagged = reinterpret_cast<int *>(&fbuf[0]);
  int nf = 0;
  for (int i = 0; i < n; ++i)
    if (cnt[i] > 0) flagged[nf++] = i;
  if (nf == 0) return NBK_OK;
  // d_packed still holds the staged, indexed bodies of the scan
  TRY(reserve(ctx, ctx->d_aux_in, (size_t)nf * sizeof(int)));
  TRY(reserve(ctx, ctx->d_out[0], (size_t)(max_pairs + 1) * sizeof(int)));
  TRY(reserve(ctx, ctx->d_out[1], (size_t)(max_pairs + 1) * sizeof(int)));
  TRY(reserve(ctx, ctx->d_out[2], (size_t)(max_pairs + 1) * sizeof(double)));
  TRY(reserve(ctx, ctx->d_scalar, 64));
  CK(cudaMemcpyAsync(ctx->d_aux_in.p, flagged, (size_t)nf * sizeof(int), cudaMemcpyHostToDevice,
                     ctx->stream));
  CK(cudaMemsetAsync(ctx->d_scalar.p, 0, 64, ctx->stream));
  binary_list_kernel<<<nf, 256, 0, ctx->stream>>>(
      (const double *)ctx->d_packed.p, n, (const int *)ctx->d_aux_in.p, eps2, threshold, max_pairs,
      (int *)ctx->d_out[0].p, (int *)ctx->d_out[1].p, (double *)ctx->d_out[2].p,
      (int *)ctx->d_scalar.p);
  CK(cudaGetLastError());
  ctx->launches++;
  int found = 0;
  CK(cudaMemcpyAsync(&found, ctx->d_scalar.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TRY(sync(ctx));
  *npairs = found;
  const int take = found < max_pairs ? found : max_pairs;
  if (take > 0) {
    CK(cudaMemcpyAsync(pi, ctx->d_out[0].p, (size_t)take * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(pj, ctx->d_out[1].p, (size_t)take * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(e, ctx->d_out[2].p, (size_t)take * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync(ctx));
    // the kernel appends in arrival order: sort into the reference's loop order (i, then j)
    std::string obuf((size_t)take * sizeof(int), '\0');
    int *ord = reinterpret_cast<int *>(&obuf[0]);
    for (int k = 0; k < take; ++k) ord[k] = k;
    std::sort(ord, ord + take, [&](int a, int b) {
      return pi[a] != pi[b] ? pi[a] < pi[b] : pj[a] < pj[b];
    });
    std::string t1((size_t)take * sizeof(int), '\0'), t2((size_t)take * sizeof(int), '\0'),
        t3((size_t)take * sizeof(double), '\0');
    int *si = reinterpret_cast<int *>(&t1[0]), *sj = reinterpret_cast<int *>(&t2[0]);
    double *se = reinterpret_cast<double *>(&t3[0]);
    for (int k = 0; k < take; ++k) {
      si[k] = pi[ord[k]];
      sj[k] = pj[ord[k]];
      se[k] = e[ord[k]];
    }
    memcpy(pi, si, (size_t)take * sizeof(int));
    memcpy(pj, sj, (size_t)take * sizeof(int));
    memcpy(e, se, (size_t)take * sizeof(double));
  }
  if (found > max_pairs)
    return fail(ctx, NBK_ERR_ARG, "nbk_binary_list: %d pairs found, room for %d", found, max_pairs);
  return NBK_OK;
}

int nbk_interact_sum(nbk_ctx *ctx, int n, int na, const double *m, const double *q, double eps2,
                     double *S) {
  if (!ctx) return NBK_ERR_ARG;
  if (n < 0 || na < 0 || na > n || !S)
    return fail(ctx, NBK_ERR_ARG, "nbk_interact_sum: bad arguments (n=%d, na=%d)", n, na);
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return NBK_OK;
  TRY(stage_bodies(ctx, n, m, q, nullptr, 1));
  TRY(reserve(ctx, ctx->d_out[0], 3 * (size_t)n * sizeof(double)));
  double *out = (double *)ctx->d_out[0].p;
  const double *packed = (const double *)ctx->d_packed.p;
  if (na == 0) {
    CK(cudaMemsetAsync(out, 0, 3 * (size_t)n * sizeof(double), ctx->stream));
  } else {
    SweepParams P = base_params(packed, eps2, 0, na, 0, n);  // active targets x everybody
    P.out[0] = out;
    TRY(K1::run(ctx, P, ctx->stream, true));
    if (n > na) {  // passive targets x active sources
      SweepParams Q = base_params(packed, eps2, na, n, 0, na);
      Q.out[0] = out + 3 * (size_t)na;
      TRY(K1::run(ctx, Q, ctx->stream, false));
    }
  }
  TRY(download(ctx, S, ctx->d_out[0], 3 * (size_t)n));
  return sync(ctx);
}

int nbk_kick_block_sum(nbk_ctx *ctx, int iend, int ibegin, const double *m, const double *q,
                       const double *v, double eta, double dt, double *dv, double *tmin) {
  if (!ctx) return NBK_ERR_ARG;
  if (iend < 0 || ibegin < 0 || ibegin > iend || !dv)
    return fail(ctx, NBK_ERR_ARG, "nbk_kick_block_sum: bad arguments (ibegin=%d, iend=%d)", ibegin,
                iend);
  CK(cudaSetDevice(ctx->device));
  if (iend == 0) return NBK_OK;
  if (tmin && !v) return fail(ctx, NBK_ERR_ARG, "time-scale needs velocities");
  const size_t n = (size_t)iend;
  TRY(stage_bodies(ctx, iend, m, q, v, 1));
  TRY(reserve(ctx, ctx->d_out[0], 3 * n * sizeof(double)));
  TRY(reserve(ctx, ctx->d_out[1], n * sizeof(double)));
  double *o0 = (double *)ctx->d_out[0].p, *o1 = (double *)ctx->d_out[1].p;
  const double *packed = (const double *)ctx->d_packed.p;
  auto sweep = [&](int t0, int t1, int s0, int s1, bool time_it) -> int {
    SweepParams P = base_params(packed, 0.0, t0, t1, s0, s1);
    P.dt = dt;
    P.eta = eta;
    P.out[0] = o0 + 3 * (size_t)t0;
    P.out[1] = o1 + t0;
    return tmin ? K4T::run(ctx, P, ctx->stream, time_it) : K4::run(ctx, P, ctx->stream, time_it);
  };
  if (ibegin == iend) {
    CK(cudaMemsetAsync(o0, 0, 3 * n * sizeof(double), ctx->stream));
    if (tmin) {
      fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(o1, n, __builtin_inf());
      CK(cudaGetLastError());
      ctx->launches++;
    }
  } else {
    TRY(sweep(ibegin, iend, 0, iend, true));                    // active x everybody
    if (ibegin > 0) TRY(sweep(0, ibegin, ibegin, iend, false));  // passive x active
  }
  TRY(download(ctx, dv, ctx->d_out[0], 3 * n));
  if (tmin) TRY(download(ctx, tmin, ctx->d_out[1], n));
  return sync(ctx);
}

int nbk_grad_v_dgrad_v_sum_predicted(nbk_ctx *ctx, int n, const double *t, const double *m,
                                     const double *q, const double *p, const double *f,
                                     const double *df, double nt, double eps2, int t0, int t1,
                                     int s0, int s1, double *gsum, double *dgsum) {
  TRY(check_ranges(ctx, n, t0, t1, s0, s1));
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return NBK_OK;
  if (!t || !m || !q || !p || !f || !df) return fail(ctx, NBK_ERR_ARG, "null body array");
  TRY(upload(ctx, ctx->d_m, m, n));
  TRY(upload(ctx, ctx->d_q, q, 3 * (size_t)n));
  TRY(upload(ctx, ctx->d_vel, p, 3 * (size_t)n));
  TRY(upload(ctx, ctx->d_t, t, n));
  TRY(upload(ctx, ctx->d_f, f, 3 * (size_t)n));
  TRY(upload(ctx, ctx->d_df, df, 3 * (size_t)n));
  TRY(reserve(ctx, ctx->d_packed, (size_t)n * PK * sizeof(double)));
  predict_pack_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(
      n, (const double *)ctx->d_t.p, (const double *)ctx->d_m.p, (const double *)ctx->d_q.p,
      (const double *)ctx->d_vel.p, (const double *)ctx->d_f.p, (const double *)ctx->d_df.p, nt,
      (double *)ctx->d_packed.p);
  CK(cudaGetLastError());
  ctx->launches++;
  return run_grad_v_dgrad_v(ctx, (const double *)ctx->d_packed.p, eps2, t0, t1, s0, s1, gsum,
                            dgsum);
}

namespace {
// Value and tangent sums over bodies already packed in ctx->d_packed and directions already
// packed in ctx->d_aux ([K][n][8]); results to the host.
int tangent_sweeps(nbk_ctx *ctx, int n, int K, double eps2, int t0, int t1, int s0, int s1,
                   double *gsum, double *dgsum, double *gsum_tan, double *dgsum_tan, bool jerk) {
  const size_t nt = (size_t)(t1 - t0);
  if (gsum || dgsum) {  // value sums from the plain kernels
    if (jerk) {
      TRY(reserve(ctx, ctx->d_out[2], 3 * nt * sizeof(double)));
      TRY(reserve(ctx, ctx->d_out[3], 3 * nt * sizeof(double)));
      if (s1 > s0) {
        SweepParams P = base_params((const double *)ctx->d_packed.p, eps2, t0, t1, s0, s1);
        P.out[0] = (double *)ctx->d_out[2].p;
        P.out[1] = (double *)ctx->d_out[3].p;
        TRY(run_k2(ctx, P, ctx->stream, false));
      } else {
        CK(cudaMemsetAsync(ctx->d_out[2].p, 0, 3 * nt * sizeof(double), ctx->stream));
        CK(cudaMemsetAsync(ctx->d_out[3].p, 0, 3 * nt * sizeof(double), ctx->stream));
      }
      if (gsum) TRY(download(ctx, gsum, ctx->d_out[2], 3 * nt));
      if (dgsum) TRY(download(ctx, dgsum, ctx->d_out[3], 3 * nt));
    } else if (gsum) {
      TRY(reserve(ctx, ctx->d_out[2], 3 * nt * sizeof(double)));
      if (s1 > s0) {
        SweepParams P = base_params((const double *)ctx->d_packed.p, eps2, t0, t1, s0, s1);
        P.out[0] = (double *)ctx->d_out[2].p;
        TRY(K1::run(ctx, P, ctx->stream, false));
      } else {
        CK(cudaMemsetAsync(ctx->d_out[2].p, 0, 3 * nt * sizeof(double), ctx->stream));
      }
      TRY(download(ctx, gsum, ctx->d_out[2], 3 * nt));
    }
  }
  if (K == 0) return sync(ctx);
  TRY(reserve(ctx, ctx->d_out[0], (size_t)K * 3 * nt * sizeof(double)));
  TRY(reserve(ctx, ctx->d_out[1], (size_t)K * 3 * nt * sizeof(double)));
  if (s1 == s0) {
    CK(cudaMemsetAsync(ctx->d_out[0].p, 0, (size_t)K * 3 * nt * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_out[1].p, 0, (size_t)K * 3 * nt * sizeof(double), ctx->stream));
  } else {
    for (int k0 = 0; k0 < K;) {
      const int kt = (K - k0 >= 3) ? 3 : 1;
      SweepParams P = base_params((const double *)ctx->d_packed.p, eps2, t0, t1, s0, s1);
      P.aux = (const double *)ctx->d_aux.p + (size_t)k0 * n * PK;
      P.aux_stride = n * PK;
      P.K = kt;
      P.out[0] = (double *)ctx->d_out[0].p + (size_t)k0 * 3 * nt;
      P.out[1] = (double *)ctx->d_out[1].p + (size_t)k0 * 3 * nt;
      if (jerk) {
        if (kt == 3) TRY((run_tangent<3, true>(ctx, P, ctx->stream)));
        else TRY((run_tangent<1, true>(ctx, P, ctx->stream)));
      } else {
        if (kt == 3) TRY((run_tangent<3, false>(ctx, P, ctx->stream)));
        else TRY((run_tangent<1, false>(ctx, P, ctx->stream)));
      }
      k0 += kt;
    }
  }
  TRY(download(ctx, gsum_tan, ctx->d_out[0], (size_t)K * 3 * nt));
  if (jerk) TRY(download(ctx, dgsum_tan, ctx->d_out[1], (size_t)K * 3 * nt));
  return sync(ctx);
}

// Shared body of the two tangent entry points.
int tangent_common(nbk_ctx *ctx, int n, const double *m, const double *q, const double *p, int K,
                   const double *dq, const double *dp, double eps2, int t0, int t1, int s0,
                   int s1, double *gsum, double *dgsum, double *gsum_tan, double *dgsum_tan,
                   bool jerk) {
  TRY(check_ranges(ctx, n, t0, t1, s0, s1));
  CK(cudaSetDevice(ctx->device));
  if (K < 0 || (K > 0 && (!dq || !gsum_tan || (jerk && (!dp || !dgsum_tan)))))
    return fail(ctx, NBK_ERR_ARG, "tangent sweep: bad direction arguments");
  if (n == 0 || t1 == t0) return NBK_OK;
  const size_t n3 = 3 * (size_t)n;
  if (!ctx->peers.empty() && s1 > s0 && K > 0 && multi_can_shard(ctx, n, s0)) {
    // one process, several GPUs: targets sharded, every device stages its own rows of the bodies
    // AND of the K direction arrays, the tangent sweeps read the other devices' rows in place
    const size_t ntg = (size_t)(t1 - t0);
    return multi_sweep(
        ctx, n, m, q, jerk ? p : nullptr, 0, eps2, t0, t1, s0, s1,
        [=](nbk_ctx *c, const Slice &sl, SweepParams P) -> int {
          const size_t cnt = (size_t)(sl.b - sl.a);
          const int ndev = 1 + (int)ctx->peers.size();
          const size_t dir = (size_t)ctx->multi_shard_rows * PK;  // doubles per direction in a shard
          int rc = reserve(c, c->d_out[0], (size_t)K * 3 * cnt * sizeof(double));
          if (rc == NBK_OK) rc = reserve(c, c->d_out[1], (size_t)K * 3 * cnt * sizeof(double));
          if (rc == NBK_OK && (gsum || dgsum)) {
            rc = reserve(c, c->d_out[2], 3 * cnt * sizeof(double));
            if (rc == NBK_OK) rc = reserve(c, c->d_out[3], 3 * cnt * sizeof(double));
            if (rc != NBK_OK) return rc;
            SweepParams V = P;
            V.out[0] = (double *)c->d_out[2].p;
            V.out[1] = (double *)c->d_out[3].p;
            rc = jerk ? run_op_peer<OpGradVDGradVB<1>>(c, V, c->stream, false)
                      : run_op_peer<OpGradV>(c, V, c->stream, false);
          }
          if (rc != NBK_OK) return rc;
          for (int k0 = 0; k0 < K && rc == NBK_OK;) {
            const int kt = (K - k0 >= 3) ? 3 : 1;
            SweepParams T = P;
            for (int e = 0; e < ndev; ++e)
              T.aux_shard[e] = (const double *)multi_dev(ctx, e)->d_aux.p + (size_t)k0 * dir;
            T.aux = T.aux_shard[sl.d] - (size_t)sl.d * ctx->multi_shard_rows * PK;
            T.aux_stride = (int)dir;
            T.K = kt;
            T.out[0] = (double *)c->d_out[0].p + (size_t)k0 * 3 * cnt;
            T.out[1] = (double *)c->d_out[1].p + (size_t)k0 * 3 * cnt;
            if (jerk) rc = kt == 3 ? run_tangent<3, true, true>(c, T, c->stream)
                                   : run_tangent<1, true, true>(c, T, c->stream);
            else rc = kt == 3 ? run_tangent<3, false, true>(c, T, c->stream)
                              : run_tangent<1, false, true>(c, T, c->stream);
            k0 += kt;
          }
          return rc;
        },
        [=](nbk_ctx *c, const Slice &sl) -> int {
          const size_t cnt = (size_t)(sl.b - sl.a), off = 3 * (size_t)(sl.a - t0);
          int rc = NBK_OK;
          if (gsum) rc = download(c, gsum + off, c->d_out[2], 3 * cnt);
          if (rc == NBK_OK && dgsum) rc = download(c, dgsum + off, c->d_out[3], 3 * cnt);
          for (int k = 0; k < K && rc == NBK_OK; ++k) {  // caller's arrays are [K][3 ntg]
            if (cudaMemcpyAsync(gsum_tan + (size_t)k * 3 * ntg + off,
                                (const double *)c->d_out[0].p + (size_t)k * 3 * cnt,
                                3 * cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
              rc = fail(c, NBK_ERR_CUDA, "cudaMemcpyAsync failed");
            if (rc == NBK_OK && jerk &&
                cudaMemcpyAsync(dgsum_tan + (size_t)k * 3 * ntg + off,
                                (const double *)c->d_out[1].p + (size_t)k * 3 * cnt,
                                3 * cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
              rc = fail(c, NBK_ERR_CUDA, "cudaMemcpyAsync failed");
          }
          return rc;
        },
        [=](nbk_ctx *c, int a, int b, int rows) -> int {
          const int cnt = b - a;
          int rc = reserve(c, c->d_aux, (size_t)K * rows * PK * sizeof(double));
          if (rc == NBK_OK) rc = reserve(c, c->d_aux_in, (size_t)K * 3 * std::max(cnt, 1) * sizeof(double));
          if (rc == NBK_OK && jerk)
            rc = reserve(c, c->d_aux_in2, (size_t)K * 3 * std::max(cnt, 1) * sizeof(double));
          for (int k = 0; k < K && rc == NBK_OK && cnt > 0; ++k) {
            double *din = (double *)c->d_aux_in.p + (size_t)k * 3 * cnt;
            double *din2 = jerk ? (double *)c->d_aux_in2.p + (size_t)k * 3 * cnt : nullptr;
            if (cudaMemcpyAsync(din, dq + (size_t)k * n3 + 3 * (size_t)a, 3 * (size_t)cnt * sizeof(double),
                                cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                (jerk && cudaMemcpyAsync(din2, dp + (size_t)k * n3 + 3 * (size_t)a,
                                         3 * (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice,
                                         c->stream) != cudaSuccess))
              rc = fail(c, NBK_ERR_CUDA, "cudaMemcpyAsync failed");
            if (rc == NBK_OK) {
              pack_aux_kernel<<<(cnt + 255) / 256, 256, 0, c->stream>>>(
                  cnt, (const double *)c->d_m.p, din, din2, 0,
                  (double *)c->d_aux.p + (size_t)k * rows * PK);
              c->launches++;
            }
          }
          return rc;
        });
  }
  TRY(stage_bodies(ctx, n, m, q, jerk ? p : nullptr, 0));
  if (K > 0) {  // directions: upload, pack into aux[K][n][8]
    TRY(reserve(ctx, ctx->d_aux, (size_t)K * n * PK * sizeof(double)));
    TRY(upload(ctx, ctx->d_aux_in, dq, (size_t)K * n3));
    if (jerk) TRY(upload(ctx, ctx->d_aux_in2, dp, (size_t)K * n3));
    for (int k = 0; k < K; ++k) {
      pack_aux_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(
          n, (const double *)ctx->d_m.p, (const double *)ctx->d_aux_in.p + (size_t)k * n3,
          jerk ? (const double *)ctx->d_aux_in2.p + (size_t)k * n3 : nullptr, 0,
          (double *)ctx->d_aux.p + (size_t)k * n * PK);
      ctx->launches++;
    }
    CK(cudaGetLastError());
  }
  return tangent_sweeps(ctx, n, K, eps2, t0, t1, s0, s1, gsum, dgsum, gsum_tan, dgsum_tan, jerk);
}
}  // namespace

// DFloat_hermite.advance_body's loop (dFloat_hermite.ml:115-135): predict every body to the dual
// time nt from its own dual state, then the tangent acc+jerk sums over the predicted bodies.
int nbk_grad_v_dgrad_v_sum_predicted_tangent(
    nbk_ctx *ctx, int n, const double *t, const double *m, const double *q, const double *p,
    const double *f, const double *df, int K, const double *t_tan, const double *q_tan,
    const double *p_tan, const double *f_tan, const double *df_tan, double nt,
    const double *nt_tan, double eps2, int t0, int t1, int s0, int s1, double *gsum,
    double *dgsum, double *gsum_tan, double *dgsum_tan) {
  TRY(check_ranges(ctx, n, t0, t1, s0, s1));
  CK(cudaSetDevice(ctx->device));
  if (!t || !m || !q || !p || !f || !df) return fail(ctx, NBK_ERR_ARG, "null body array");
  if (K < 0 || (K > 0 && (!q_tan || !p_tan || !f_tan || !df_tan || !gsum_tan || !dgsum_tan)))
    return fail(ctx, NBK_ERR_ARG, "predicted tangent sweep: bad direction arguments");
  if (n == 0 || t1 == t0) return NBK_OK;
  const size_t n3 = 3 * (size_t)n, Kn3 = (size_t)K * n3;
  TRY(upload(ctx, ctx->d_m, m, n));
  TRY(upload(ctx, ctx->d_q, q, n3));
  TRY(upload(ctx, ctx->d_vel, p, n3));
  TRY(upload(ctx, ctx->d_t, t, n));
  TRY(upload(ctx, ctx->d_f, f, n3));
  TRY(upload(ctx, ctx->d_df, df, n3));
  TRY(reserve(ctx, ctx->d_packed, (size_t)n * PK * sizeof(double)));
  predict_pack_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(
      n, (const double *)ctx->d_t.p, (const double *)ctx->d_m.p, (const double *)ctx->d_q.p,
      (const double *)ctx->d_vel.p, (const double *)ctx->d_f.p, (const double *)ctx->d_df.p, nt,
      (double *)ctx->d_packed.p);
  CK(cudaGetLastError());
  ctx->launches++;
  if (K > 0) {
    // tangent inputs in one buffer: [q_tan | p_tan | f_tan | df_tan] (each [K][3n]), then
    // t_tan [K][n] (zeros if NULL) and nt_tan [K] (zeros if NULL)
    TRY(reserve(ctx, ctx->d_aux_in, (4 * Kn3 + (size_t)K * n + K) * sizeof(double)));
    double *base = (double *)ctx->d_aux_in.p;
    double *d_tt = base + 4 * Kn3, *d_ntt = d_tt + (size_t)K * n;
    const double *src[4] = {q_tan, p_tan, f_tan, df_tan};
    for (int a = 0; a < 4; ++a)
      CK(cudaMemcpyAsync(base + a * Kn3, src[a], Kn3 * sizeof(double), cudaMemcpyHostToDevice,
                         ctx->stream));
    if (t_tan)
      CK(cudaMemcpyAsync(d_tt, t_tan, (size_t)K * n * sizeof(double), cudaMemcpyHostToDevice,
                         ctx->stream));
    else
      CK(cudaMemsetAsync(d_tt, 0, (size_t)K * n * sizeof(double), ctx->stream));
    if (nt_tan)
      CK(cudaMemcpyAsync(d_ntt, nt_tan, (size_t)K * sizeof(double), cudaMemcpyHostToDevice,
                         ctx->stream));
    else
      CK(cudaMemsetAsync(d_ntt, 0, (size_t)K * sizeof(double), ctx->stream));
    TRY(reserve(ctx, ctx->d_aux, (size_t)K * n * PK * sizeof(double)));
    dim3 grid((n + 255) / 256, K);
    predict_pack_tangent_kernel<<<grid, 256, 0, ctx->stream>>>(
        n, (const double *)ctx->d_t.p, (const double *)ctx->d_m.p, (const double *)ctx->d_q.p,
        (const double *)ctx->d_vel.p, (const double *)ctx->d_f.p, (const double *)ctx->d_df.p, nt,
        d_tt, base, base + Kn3, base + 2 * Kn3, base + 3 * Kn3, d_ntt, (double *)ctx->d_aux.p,
        (size_t)n * PK);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  return tangent_sweeps(ctx, n, K, eps2, t0, t1, s0, s1, gsum, dgsum, gsum_tan, dgsum_tan, true);
}

int nbk_grad_v_dgrad_v_sum_tangent(nbk_ctx *ctx, int n, const double *m, const double *q,
                                   const double *p, int K, const double *dq, const double *dp,
                                   double eps2, int t0, int t1, int s0, int s1, double *gsum,
                                   double *dgsum, double *gsum_tan, double *dgsum_tan) {
  if (!ctx) return NBK_ERR_ARG;
  if (!p) return fail(ctx, NBK_ERR_ARG, "null momentum array");
  return tangent_common(ctx, n, m, q, p, K, dq, dp, eps2, t0, t1, s0, s1, gsum, dgsum, gsum_tan,
                        dgsum_tan, true);
}

int nbk_grad_v_sum_tangent(nbk_ctx *ctx, int n, const double *m, const double *q, int K,
                           const double *dq, double eps2, int t0, int t1, int s0, int s1,
                           double *gsum, double *gsum_tan) {
  if (!ctx) return NBK_ERR_ARG;
  return tangent_common(ctx, n, m, q, nullptr, K, dq, nullptr, eps2, t0, t1, s0, s1, gsum,
                        nullptr, gsum_tan, nullptr, false);
}

// ---- device-resident bodies --------------------------------------------------------------------
int nbk_bodies_upload(nbk_ctx *ctx, int n, const double *m, const double *q, const double *vel,
                      int is_velocity) {
  if (!ctx) return NBK_ERR_ARG;
  if (n < 0 || !m || !q) return fail(ctx, NBK_ERR_ARG, "nbk_bodies_upload: bad arguments");
  CK(cudaSetDevice(ctx->device));
  TRY(upload(ctx, ctx->d_res_m, m, n));
  TRY(upload(ctx, ctx->d_q, q, 3 * (size_t)n));
  if (vel) TRY(upload(ctx, ctx->d_vel, vel, 3 * (size_t)n));
  TRY(reserve(ctx, ctx->d_res_packed, (size_t)n * PK * sizeof(double)));
  TRY(pack(ctx, n, (const double *)ctx->d_res_m.p, (const double *)ctx->d_q.p,
           vel ? (const double *)ctx->d_vel.p : nullptr, is_velocity,
           (double *)ctx->d_res_packed.p, ctx->stream));
  TRY(sync(ctx));
  ctx->resident_n = n;
  return NBK_OK;
}

int nbk_bodies_update(nbk_ctx *ctx, int i0, int i1, const double *q, const double *vel,
                      int is_velocity) {
  if (!ctx) return NBK_ERR_ARG;
  if (i0 < 0 || i1 < i0 || i1 > ctx->resident_n || !q)
    return fail(ctx, NBK_ERR_ARG, "nbk_bodies_update: bad range [%d,%d) of %d", i0, i1,
                ctx->resident_n);
  CK(cudaSetDevice(ctx->device));
  const int cnt = i1 - i0;
  if (cnt == 0) return NBK_OK;
  TRY(upload(ctx, ctx->d_q, q, 3 * (size_t)cnt));
  if (vel) TRY(upload(ctx, ctx->d_vel, vel, 3 * (size_t)cnt));
  TRY(pack(ctx, cnt, (const double *)ctx->d_res_m.p + i0, (const double *)ctx->d_q.p,
           vel ? (const double *)ctx->d_vel.p : nullptr, is_velocity,
           (double *)ctx->d_res_packed.p + (size_t)i0 * PK, ctx->stream));
  return sync(ctx);
}

int nbk_resident_count(const nbk_ctx *ctx) { return ctx ? ctx->resident_n : 0; }

#define NEED_RESIDENT()                                                            \
  do {                                                                             \
    if (!ctx) return NBK_ERR_ARG;                                                  \
    if (ctx->resident_n == 0)                                                      \
      return fail(ctx, NBK_ERR_STATE, "no resident bodies: call nbk_bodies_upload"); \
    TRY(check_ranges(ctx, ctx->resident_n, t0, t1, s0, s1));                       \
    CK(cudaSetDevice(ctx->device));                                                \
  } while (0)

int nbk_resident_grad_v_sum(nbk_ctx *ctx, double eps2, int t0, int t1, int s0, int s1,
                            double *gsum) {
  NEED_RESIDENT();
  return run_grad_v(ctx, (const double *)ctx->d_res_packed.p, eps2, t0, t1, s0, s1, gsum);
}

int nbk_resident_grad_v_dgrad_v_sum(nbk_ctx *ctx, double eps2, int t0, int t1, int s0, int s1,
                                    double *gsum, double *dgsum) {
  NEED_RESIDENT();
  return run_grad_v_dgrad_v(ctx, (const double *)ctx->d_res_packed.p, eps2, t0, t1, s0, s1, gsum,
                            dgsum);
}

int nbk_resident_v_sum(nbk_ctx *ctx, double eps2, int t0, int t1, int s0, int s1, double *phi,
                       double *total) {
  NEED_RESIDENT();
  return run_v(ctx, (const double *)ctx->d_res_packed.p, eps2, t0, t1, s0, s1, phi, total);
}

int nbk_resident_kick_sum(nbk_ctx *ctx, double eta, double dt, int t0, int t1, int s0, int s1,
                          double *dv, double *tmin) {
  NEED_RESIDENT();
  return run_kick(ctx, (const double *)ctx->d_res_packed.p, eta, dt, t0, t1, s0, s1, dv, tmin);
}

int nbk_resident_leapfrog_steps(nbk_ctx *ctx, double dt, int nsteps, double *q, double *v) {
  if (!ctx) return NBK_ERR_ARG;
  const int n = ctx->resident_n;
  if (n == 0) return fail(ctx, NBK_ERR_STATE, "no resident bodies: call nbk_bodies_upload");
  if (nsteps < 0) return fail(ctx, NBK_ERR_ARG, "nbk_resident_leapfrog_steps: nsteps < 0");
  CK(cudaSetDevice(ctx->device));
  double *packed = (double *)ctx->d_res_packed.p;
  TRY(reserve(ctx, ctx->d_out[0], 3 * (size_t)n * sizeof(double)));
  const double dt2 = dt / 2.0;  // leapfrog.ml:121
  const int nb = (n + 255) / 256;
  for (int s = 0; s < nsteps; ++s) {
    leapfrog_drift_kernel<<<nb, 256, 0, ctx->stream>>>(packed, n, dt2);
    SweepParams P = base_params(packed, 0.0, 0, n, 0, n);
    P.dt = dt;
    P.out[0] = (double *)ctx->d_out[0].p;
    TRY(K4::run(ctx, P, ctx->stream, s == nsteps - 1));
    leapfrog_apply_kick_kernel<<<nb, 256, 0, ctx->stream>>>(packed, (const double *)ctx->d_out[0].p, n);
    leapfrog_drift_kernel<<<nb, 256, 0, ctx->stream>>>(packed, n, dt2);
    CK(cudaGetLastError());
    ctx->launches += 3;
  }
  if (q || v) {
    TRY(reserve(ctx, ctx->d_q, 3 * (size_t)n * sizeof(double)));
    TRY(reserve(ctx, ctx->d_vel, 3 * (size_t)n * sizeof(double)));
    unpack_kernel<<<nb, 256, 0, ctx->stream>>>(packed, n, (double *)ctx->d_q.p,
                                               (double *)ctx->d_vel.p);
    CK(cudaGetLastError());
    ctx->launches++;
    if (q) TRY(download(ctx, q, ctx->d_q, 3 * (size_t)n));
    if (v) TRY(download(ctx, v, ctx->d_vel, 3 * (size_t)n));
  }
  return sync(ctx);
}

// ---- device-resident Advancer.A records ----------------------------------------------------------
#define NEED_ADV()                                                                          \
  do {                                                                                      \
    if (!ctx) return NBK_ERR_ARG;                                                           \
    if (ctx->adv_n == 0)                                                                    \
      return fail(ctx, NBK_ERR_STATE, "no resident Advancer.A records: call nbk_adv_upload"); \
    CK(cudaSetDevice(ctx->device));                                                         \
  } while (0)

int nbk_adv_upload(nbk_ctx *ctx, int n, const double *records) {
  if (!ctx) return NBK_ERR_ARG;
  if (n <= 0 || !records) return fail(ctx, NBK_ERR_ARG, "nbk_adv_upload: bad arguments");
  CK(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)n * AB * sizeof(double);
  TRY(reserve(ctx, ctx->d_adv, bytes));
  TRY(reserve(ctx, ctx->d_adv2, bytes));
  CK(cudaMemcpyAsync(ctx->d_adv.p, records, bytes, cudaMemcpyHostToDevice, ctx->stream));
  TRY(sync(ctx));
  ctx->adv_n = n;
  return NBK_OK;
}

int nbk_adv_download(nbk_ctx *ctx, int i0, int i1, double *records) {
  NEED_ADV();
  if (i0 < 0 || i1 < i0 || i1 > ctx->adv_n || !records)
    return fail(ctx, NBK_ERR_ARG, "nbk_adv_download: bad range [%d,%d)", i0, i1);
  if (i1 == i0) return NBK_OK;
  CK(cudaMemcpyAsync(records, (const double *)ctx->d_adv.p + (size_t)i0 * AB,
                     (size_t)(i1 - i0) * AB * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}

int nbk_adv_permute(nbk_ctx *ctx, int count, const int *perm) {
  NEED_ADV();
  if (count < 0 || count > ctx->adv_n || (count > 0 && !perm))
    return fail(ctx, NBK_ERR_ARG, "nbk_adv_permute: bad arguments");
  if (count == 0) return NBK_OK;
  TRY(reserve(ctx, ctx->d_aux_in, (size_t)count * sizeof(int)));
  CK(cudaMemcpyAsync(ctx->d_aux_in.p, perm, (size_t)count * sizeof(int), cudaMemcpyHostToDevice,
                     ctx->stream));
  const size_t elems = (size_t)count * AB;
  adv_gather_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, ctx->stream>>>(
      (const double *)ctx->d_adv.p, (double *)ctx->d_adv2.p, (const int *)ctx->d_aux_in.p, count);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->d_adv.p, ctx->d_adv2.p, elems * sizeof(double), cudaMemcpyDeviceToDevice,
                     ctx->stream));
  ctx->launches++;
  // perm lives in caller memory: make sure the copy has been consumed before returning
  return sync(ctx);
}

int nbk_adv_begin_step(nbk_ctx *ctx, int imin, int imax, double hnew) {
  NEED_ADV();
  if (imin < 0 || imax < imin || imax > ctx->adv_n)
    return fail(ctx, NBK_ERR_ARG, "nbk_adv_begin_step: bad range [%d,%d)", imin, imax);
  if (imax == imin) return NBK_OK;
  adv_begin_kernel<<<(imax - imin + 127) / 128, 128, 0, ctx->stream>>>((double *)ctx->d_adv.p, imin,
                                                                        imax, hnew);
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;
}

int nbk_adv_interact(nbk_ctx *ctx, int imin, int imax, double t, double vC, double eps2) {
  NEED_ADV();
  const int n = ctx->adv_n;
  if (imin < 0 || imax < imin || imax > n)
    return fail(ctx, NBK_ERR_ARG, "nbk_adv_interact: bad range [%d,%d)", imin, imax);
  const int nn = n - imin, na = imax - imin;
  if (nn == 0) return NBK_OK;
  TRY(reserve(ctx, ctx->d_packed, (size_t)nn * PK * sizeof(double)));
  TRY(reserve(ctx, ctx->d_out[0], 3 * (size_t)nn * sizeof(double)));
  double *packed = (double *)ctx->d_packed.p, *S = (double *)ctx->d_out[0].p;
  adv_predict_pack_kernel<<<(nn + 127) / 128, 128, 0, ctx->stream>>>((double *)ctx->d_adv.p, imin, n,
                                                                      t, packed);
  CK(cudaGetLastError());
  ctx->launches++;
  if (na == 0) return NBK_OK;
  SweepParams P = base_params(packed, eps2, 0, na, 0, nn);  // active x everybody
  P.out[0] = S;
  TRY(K1::run(ctx, P, ctx->stream, false));
  if (nn > na) {  // passive x active
    SweepParams Q = base_params(packed, eps2, na, nn, 0, na);
    Q.out[0] = S + 3 * (size_t)na;
    TRY(K1::run(ctx, Q, ctx->stream, false));
  }
  adv_accumulate_kernel<<<(nn + 127) / 128, 128, 0, ctx->stream>>>((double *)ctx->d_adv.p, imin, n,
                                                                    vC, S);
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;  // enqueued only: the next read-back synchronises
}

int nbk_adv_finish(nbk_ctx *ctx, int imin, int imax, double t, double *records) {
  NEED_ADV();
  if (imin < 0 || imax < imin || imax > ctx->adv_n)
    return fail(ctx, NBK_ERR_ARG, "nbk_adv_finish: bad range [%d,%d)", imin, imax);
  if (imax == imin) return NBK_OK;
  adv_finish_kernel<<<(imax - imin + 127) / 128, 128, 0, ctx->stream>>>((double *)ctx->d_adv.p, imin,
                                                                         imax, t);
  CK(cudaGetLastError());
  ctx->launches++;
  if (records) return nbk_adv_download(ctx, imin, imax, records);
  return NBK_OK;
}

int nbk_adv_get_t(nbk_ctx *ctx, int i, double *t) {
  NEED_ADV();
  if (i < 0 || i >= ctx->adv_n || !t) return fail(ctx, NBK_ERR_ARG, "nbk_adv_get_t: bad index %d", i);
  CK(cudaMemcpyAsync(ctx->h_scalar, (const double *)ctx->d_adv.p + (size_t)i * AB, sizeof(double),
                     cudaMemcpyDeviceToHost, ctx->stream));
  TRY(sync(ctx));
  *t = ctx->h_scalar[0];
  return NBK_OK;
}

int nbk_adv_init_first(nbk_ctx *ctx, double eps2) {
  NEED_ADV();
  const int n = ctx->adv_n;
  TRY(reserve(ctx, ctx->d_packed, (size_t)n * PK * sizeof(double)));
  TRY(reserve(ctx, ctx->d_out[0], 3 * (size_t)n * sizeof(double)));
  TRY(reserve(ctx, ctx->d_out[1], 3 * (size_t)n * sizeof(double)));
  adv_pack_full_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((const double *)ctx->d_adv.p, n,
                                                                  (double *)ctx->d_packed.p);
  CK(cudaGetLastError());
  ctx->launches++;
  if (n > 1) {
    SweepParams P = base_params((const double *)ctx->d_packed.p, eps2, 0, n, 0, n);
    P.out[0] = (double *)ctx->d_out[0].p;
    P.out[1] = (double *)ctx->d_out[1].p;
    TRY(run_k2(ctx, P, ctx->stream, false));
  } else {
    CK(cudaMemsetAsync(ctx->d_out[0].p, 0, 3 * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_out[1].p, 0, 3 * sizeof(double), ctx->stream));
  }
  adv_init_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(
      (double *)ctx->d_adv.p, n, (const double *)ctx->d_out[0].p, (const double *)ctx->d_out[1].p);
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;
}

int nbk_adv_range_max(void) { return ADT_MAX; }

int nbk_adv_advance_range(nbk_ctx *ctx, int imax, double dt, double sf, double eps2, double *hmax,
                          int *perm, long long *stats) {
  NEED_ADV();
  const int n = ctx->adv_n;
  if (imax < 1 || imax > n || imax > ADT_MAX || !hmax || !perm || !(dt > 0.0))
    return fail(ctx, NBK_ERR_ARG, "nbk_adv_advance_range: bad arguments (imax=%d of %d, limit %d, dt=%g)",
                imax, n, ADT_MAX, dt);
  // CTA size by n: one slot per thread wherever the chip is wide enough, few resident warps else
  // (512 threads leave 128 registers per thread and spill; two slots per thread at 256 measured faster)
  int T = n <= ctx->sm_count * 128 ? 128 : 256;
  int G = std::min(ctx->sm_count, (n + T - 1) / T);
  if (const char *e = getenv("NBK_ADT_SHAPE")) {  // "threads,ctas": a knob for tests and profiles
    int t = 0, g = 0;
    if (sscanf(e, "%d,%d", &t, &g) == 2 && (t == 128 || t == 256 || t == 512) && g >= 1 &&
        g <= ctx->sm_count) {
      T = t;
      G = g;
    }
  }
  // workspace: column-major records [AB][ld] | tact [2][ADT_MAX][6] | partial [2][G][ADT_MAX][3]
  // | hmax, hmax_out [ADT_MAX] each | out [16] | barrier | map [ADT_MAX] ints
  const size_t ld = ((size_t)n + 31) & ~(size_t)31;
  const size_t o_tact = (size_t)AB * ld, o_part = o_tact + 2 * (size_t)ADT_MAX * 6,
               o_hmax = o_part + 2 * (size_t)G * ADT_MAX * 3, o_hout = o_hmax + ADT_MAX,
               o_out = o_hout + ADT_MAX, o_bar = o_out + 16, o_map = o_bar + 2,
               total = o_map + ADT_MAX / 2;
  TRY(reserve(ctx, ctx->d_adt, total * sizeof(double)));
  double *w = (double *)ctx->d_adt.p;
  cudaStream_t st = ctx->stream;
  if (!ctx->adt_attr_set) {
    CK(cudaFuncSetAttribute(adv_tree_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)ADT_SMEM));
    CK(cudaFuncSetAttribute(adv_tree_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)ADT_SMEM));
    ctx->adt_attr_set = true;
  }
  CK(cudaMemcpyAsync(w + o_hmax, hmax, (size_t)imax * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(w + o_out, 0, 18 * sizeof(double), st));  // out[16] and the barrier counter
  AdvTreeParams P;
  P.rec = (double *)ctx->d_adv.p;
  P.soa = w;
  P.ld = ld;
  P.tact = w + o_tact;
  P.partial = w + o_part;
  P.hmax = w + o_hmax;
  P.hmax_out = w + o_hout;
  P.out = (long long *)(w + o_out);
  P.bar = (unsigned long long *)(w + o_bar);
  P.map = (int *)(w + o_map);
  P.n = n;
  P.imax = imax;
  P.dt = dt;
  P.sf = sf;
  P.eps2 = eps2;
  P.cluster = (G <= 8 && !getenv("NBK_ADT_NOCLUSTER")) ? 1 : 0;
  void *args[] = {&P};
  const void *kern = T <= 256 ? (const void *)adv_tree_kernel<256> : (const void *)adv_tree_kernel<512>;
  if (P.cluster) {  // the whole grid as one thread-block cluster: hardware barrier between the phases
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G);
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = ADT_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelExC(&cfg, kern, args));
  } else {
    CK(cudaLaunchCooperativeKernel(kern, dim3(G), dim3(T), args, ADT_SMEM, st));
  }
  ctx->launches++;  // the kernel also writes the rows back in their new order (sort_range's moves)
  long long out[16] = {0};
  CK(cudaMemcpyAsync(hmax, P.hmax_out, (size_t)imax * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(perm, P.map, (size_t)imax * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(out, P.out, sizeof out, cudaMemcpyDeviceToHost, st));
  TRY(sync(ctx));
  if (out[2] == -1)
    return fail(ctx, NBK_ERR_STATE, "nbk_adv_advance_range: block recursion deeper than %d frames",
                ADT_DEPTH);
  if (out[2] != 0)
    return fail(ctx, NBK_ERR_STATE,
                "nbk_adv_advance_range: empty block at the end of the array (index out of bounds)");
  if (stats) {
    stats[0] += out[0];
    stats[1] += out[1];
  }
  for (int k = 0; k < 10; ++k) ctx->adt_prof[k] += out[3 + k];
  ctx->adt_prof[10] += 1;
  return NBK_OK;
}

int nbk_adv_range_profile(nbk_ctx *ctx, long long *prof, int reset) {
  if (!ctx || !prof) return NBK_ERR_ARG;
  for (int k = 0; k < 11; ++k) prof[k] = ctx->adt_prof[k];
  if (reset) for (int k = 0; k < 11; ++k) ctx->adt_prof[k] = 0;
  return NBK_OK;
}

int nbk_adv_energy(nbk_ctx *ctx, double eps2, double *ke, double *pe) {
  NEED_ADV();
  const int n = ctx->adv_n;
  // Energy.energy needs only (m, q, p): taken from the resident records, nothing leaves the
  // device but the two sums.  Same kernels and summation trees as nbk_energy.
  TRY(reserve(ctx, ctx->d_packed, (size_t)n * PK * sizeof(double)));
  TRY(reserve(ctx, ctx->d_out[0], (size_t)n * sizeof(double)));
  TRY(reserve(ctx, ctx->d_scalar, 64));
  const double *rec = (const double *)ctx->d_adv.p;
  double *o = (double *)ctx->d_out[0].p, *sc = (double *)ctx->d_scalar.p;
  cudaStream_t st = ctx->stream;
  ctx->h_scalar[0] = ctx->h_scalar[1] = 0.0;
  if (pe) {
    adv_pack_full_kernel<<<(n + 255) / 256, 256, 0, st>>>(rec, n, (double *)ctx->d_packed.p);
    CK(cudaGetLastError());
    ctx->launches++;
    if (n > 1) {
      SweepParams P = base_params((const double *)ctx->d_packed.p, eps2, 0, n, 0, n);
      P.out[0] = o;
      TRY(K3::run(ctx, P, st, true));
    } else {
      CK(cudaMemsetAsync(o, 0, (size_t)n * sizeof(double), st));
    }
    sum_kernel<<<1, 1024, 0, st>>>(o, n, sc);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  if (ke) {
    adv_kinetic_kernel<<<(n + 255) / 256, 256, 0, st>>>(rec, n, o);
    sum_kernel<<<1, 1024, 0, st>>>(o, n, sc + 1);
    CK(cudaGetLastError());
    ctx->launches += 2;
  }
  CK(cudaMemcpyAsync(ctx->h_scalar, sc, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TRY(sync(ctx));
  if (pe) *pe = 0.5 * ctx->h_scalar[0];
  if (ke) *ke = ctx->h_scalar[1];
  return NBK_OK;
}

// ---- device-resident Leapfrog.b array ------------------------------------------------------------
#define NEED_LF()                                                                          \
  do {                                                                                     \
    if (!ctx) return NBK_ERR_ARG;                                                          \
    if (ctx->lf_n == 0)                                                                    \
      return fail(ctx, NBK_ERR_STATE, "no resident Leapfrog bodies: call nbk_lf_upload");   \
    CK(cudaSetDevice(ctx->device));                                                        \
  } while (0)

int nbk_lf_upload(nbk_ctx *ctx, int n, const double *dt_max, const double *m, const double *q,
                  const double *v) {
  if (!ctx) return NBK_ERR_ARG;
  if (n <= 0 || !dt_max || !m || !q || !v) return fail(ctx, NBK_ERR_ARG, "nbk_lf_upload: bad arguments");
  CK(cudaSetDevice(ctx->device));
  TRY(upload(ctx, ctx->d_m, m, n));
  TRY(upload(ctx, ctx->d_q, q, 3 * (size_t)n));
  TRY(upload(ctx, ctx->d_vel, v, 3 * (size_t)n));
  TRY(upload(ctx, ctx->d_lf_dtmax, dt_max, n));
  TRY(reserve(ctx, ctx->d_lf_dtmax2, (size_t)n * sizeof(double)));
  TRY(reserve(ctx, ctx->d_lf_packed, (size_t)n * PK * sizeof(double)));
  TRY(reserve(ctx, ctx->d_lf_packed2, (size_t)n * PK * sizeof(double)));
  TRY(pack(ctx, n, (const double *)ctx->d_m.p, (const double *)ctx->d_q.p,
           (const double *)ctx->d_vel.p, 1, (double *)ctx->d_lf_packed.p, ctx->stream));
  if (ctx->h_lf_cap < (size_t)n) {  // pinned: n doubles for dt_max, then n ints for a permutation
    if (ctx->h_lf_stage) CK(cudaFreeHost(ctx->h_lf_stage));
    ctx->h_lf_stage = nullptr;
    ctx->h_lf_cap = 0;
    CK(cudaMallocHost((void **)&ctx->h_lf_stage, (size_t)n * (sizeof(double) + sizeof(int))));
    ctx->h_lf_cap = (size_t)n;
  }
  if (!ctx->ev_lf) CK(cudaEventCreateWithFlags(&ctx->ev_lf, cudaEventDisableTiming));
  CK(cudaEventRecord(ctx->ev_lf, ctx->stream));
  ctx->lf_n = n;
  return sync(ctx);
}

int nbk_lf_download(nbk_ctx *ctx, double *dt_max, double *q, double *v) {
  NEED_LF();
  const int n = ctx->lf_n;
  if (q || v) {
    TRY(reserve(ctx, ctx->d_q, 3 * (size_t)n * sizeof(double)));
    TRY(reserve(ctx, ctx->d_vel, 3 * (size_t)n * sizeof(double)));
    unpack_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((const double *)ctx->d_lf_packed.p, n,
                                                           (double *)ctx->d_q.p, (double *)ctx->d_vel.p);
    CK(cudaGetLastError());
    ctx->launches++;
    if (q) TRY(download(ctx, q, ctx->d_q, 3 * (size_t)n));
    if (v) TRY(download(ctx, v, ctx->d_vel, 3 * (size_t)n));
  }
  if (dt_max) TRY(download(ctx, dt_max, ctx->d_lf_dtmax, n));
  return sync(ctx);
}

int nbk_lf_permute(nbk_ctx *ctx, int count, const int *perm) {
  NEED_LF();
  if (count < 0 || count > ctx->lf_n || (count > 0 && !perm))
    return fail(ctx, NBK_ERR_ARG, "nbk_lf_permute: bad arguments");
  if (count == 0) return NBK_OK;
  TRY(reserve(ctx, ctx->d_aux_in, (size_t)count * sizeof(int)));
  // stage the permutation in pinned memory so the call need not wait for the copy
  int *stage = reinterpret_cast<int *>(ctx->h_lf_stage + ctx->h_lf_cap);
  CK(cudaEventSynchronize(ctx->ev_lf));  // the previous permutation has left the buffer
  memcpy(stage, perm, (size_t)count * sizeof(int));
  CK(cudaMemcpyAsync(ctx->d_aux_in.p, stage, (size_t)count * sizeof(int), cudaMemcpyHostToDevice,
                     ctx->stream));
  CK(cudaEventRecord(ctx->ev_lf, ctx->stream));
  const size_t elems = (size_t)count * PK;
  lf_gather_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, ctx->stream>>>(
      (const double *)ctx->d_lf_packed.p, (double *)ctx->d_lf_packed2.p,
      (const double *)ctx->d_lf_dtmax.p, (double *)ctx->d_lf_dtmax2.p, (const int *)ctx->d_aux_in.p,
      count);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaMemcpyAsync(ctx->d_lf_packed.p, ctx->d_lf_packed2.p, elems * sizeof(double),
                     cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_lf_dtmax.p, ctx->d_lf_dtmax2.p, (size_t)count * sizeof(double),
                     cudaMemcpyDeviceToDevice, ctx->stream));
  return NBK_OK;
}

int nbk_lf_sort_sub(nbk_ctx *ctx, int iend, int *perm) {
  NEED_LF();
  if (iend < 0 || iend > ctx->lf_n || (iend > 0 && !perm))
    return fail(ctx, NBK_ERR_ARG, "nbk_lf_sort_sub: bad arguments");
  if (iend == 0) return NBK_OK;
  TRY(reserve(ctx, ctx->d_aux_in, (size_t)iend * sizeof(int)));
  TRY(reserve(ctx, ctx->d_aux_in2, (size_t)iend * (sizeof(unsigned long long) + sizeof(int))));
  unsigned long long *ukey = (unsigned long long *)ctx->d_aux_in2.p;
  int *rank = (int *)(ukey + iend);
  lf_sort_key_kernel<<<(iend + 255) / 256, 256, 0, ctx->stream>>>((const double *)ctx->d_lf_dtmax.p,
                                                                 iend, ukey, rank);
  dim3 rgrid((iend + 255) / 256, (iend + LFS_CHUNK - 1) / LFS_CHUNK);
  lf_rank_kernel<<<rgrid, 256, 0, ctx->stream>>>(ukey, iend, rank);
  lf_scatter_perm_kernel<<<(iend + 255) / 256, 256, 0, ctx->stream>>>(rank, iend,
                                                                     (int *)ctx->d_aux_in.p);
  CK(cudaGetLastError());
  ctx->launches += 2;
  const size_t elems = (size_t)iend * PK;
  lf_gather_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, ctx->stream>>>(
      (const double *)ctx->d_lf_packed.p, (double *)ctx->d_lf_packed2.p,
      (const double *)ctx->d_lf_dtmax.p, (double *)ctx->d_lf_dtmax2.p, (const int *)ctx->d_aux_in.p,
      iend);
  CK(cudaGetLastError());
  ctx->launches += 2;
  CK(cudaMemcpyAsync(ctx->d_lf_packed.p, ctx->d_lf_packed2.p, elems * sizeof(double),
                     cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_lf_dtmax.p, ctx->d_lf_dtmax2.p, (size_t)iend * sizeof(double),
                     cudaMemcpyDeviceToDevice, ctx->stream));
  // the permutation goes home through the pinned staging buffer (shared with nbk_lf_permute)
  int *stage = reinterpret_cast<int *>(ctx->h_lf_stage + ctx->h_lf_cap);
  CK(cudaEventSynchronize(ctx->ev_lf));
  CK(cudaMemcpyAsync(stage, ctx->d_aux_in.p, (size_t)iend * sizeof(int), cudaMemcpyDeviceToHost,
                     ctx->stream));
  TRY(sync(ctx));
  memcpy(perm, stage, (size_t)iend * sizeof(int));
  return NBK_OK;
}

int nbk_lf_begin(nbk_ctx *ctx, int ibegin, int iend) {
  NEED_LF();
  if (ibegin < 0 || iend < ibegin || iend > ctx->lf_n)
    return fail(ctx, NBK_ERR_ARG, "nbk_lf_begin: bad range [%d,%d)", ibegin, iend);
  if (iend == ibegin) return NBK_OK;
  const size_t cnt = (size_t)(iend - ibegin);
  fill_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(
      (double *)ctx->d_lf_dtmax.p + ibegin, cnt, __builtin_inf());
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;
}

int nbk_lf_drift(nbk_ctx *ctx, int i0, int i1, double dt) {
  NEED_LF();
  if (i0 < 0 || i1 < i0 || i1 > ctx->lf_n)
    return fail(ctx, NBK_ERR_ARG, "nbk_lf_drift: bad range [%d,%d)", i0, i1);
  if (i1 == i0) return NBK_OK;
  lf_drift_kernel<<<(i1 - i0 + 255) / 256, 256, 0, ctx->stream>>>((double *)ctx->d_lf_packed.p, i0, i1, dt);
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;
}

int nbk_lf_kick_block(nbk_ctx *ctx, int iend, int ibegin, double eta, double dt, int with_tscale,
                      int with_drift, double drift_after, double *dt_max) {
  NEED_LF();
  if (iend < 0 || iend > ctx->lf_n || ibegin < 0 || ibegin > iend)
    return fail(ctx, NBK_ERR_ARG, "nbk_lf_kick_block: bad arguments (ibegin=%d, iend=%d)", ibegin, iend);
  if (iend == 0 || ibegin == iend) return NBK_OK;
  const size_t n = (size_t)iend;
  TRY(reserve(ctx, ctx->d_out[0], 3 * n * sizeof(double)));
  TRY(reserve(ctx, ctx->d_out[1], n * sizeof(double)));
  double *o0 = (double *)ctx->d_out[0].p, *o1 = (double *)ctx->d_out[1].p;
  const double *packed = (const double *)ctx->d_lf_packed.p;
  auto sweep = [&](int t0, int t1, int s0, int s1) -> int {
    SweepParams P = base_params(packed, 0.0, t0, t1, s0, s1);
    P.dt = dt;
    P.eta = eta;
    P.out[0] = o0 + 3 * (size_t)t0;
    P.out[1] = o1 + t0;
    return with_tscale ? K4T::run(ctx, P, ctx->stream, false) : K4::run(ctx, P, ctx->stream, false);
  };
  TRY(sweep(ibegin, iend, 0, iend));                    // active x everybody
  if (ibegin > 0) TRY(sweep(0, ibegin, ibegin, iend));  // passive x active
  lf_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
      (double *)ctx->d_lf_packed.p, (double *)ctx->d_lf_dtmax.p, o0, with_tscale ? o1 : nullptr,
      ibegin, iend, with_drift, drift_after);
  CK(cudaGetLastError());
  ctx->launches++;
  if (with_tscale && dt_max) {
    CK(cudaMemcpyAsync(ctx->h_lf_stage, ctx->d_lf_dtmax.p, n * sizeof(double),
                       cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync(ctx));
    memcpy(dt_max, ctx->h_lf_stage, n * sizeof(double));
  }
  return NBK_OK;
}

int nbk_lf_subsystem_max(void) { return LFT_MAX; }

int nbk_lf_advance_subsystem(nbk_ctx *ctx, int iend, double dt, double eta, double *t, double *dt_max,
                             int *perm, long long *nkicks) {
  NEED_LF();
  if (iend < 1 || iend > ctx->lf_n || iend > LFT_MAX || !t || !dt_max || !perm || !(dt >= 0.0))
    return fail(ctx, NBK_ERR_ARG, "nbk_lf_advance_subsystem: bad arguments (iend=%d of %d, limit %d, dt=%g)",
                iend, ctx->lf_n, LFT_MAX, dt);
  // workspace: t [LFT_MAX] | out [2] | map [LFT_MAX] ints
  TRY(reserve(ctx, ctx->d_lft, (size_t)(LFT_MAX + 2 + LFT_MAX / 2) * sizeof(double)));
  double *w = (double *)ctx->d_lft.p;
  cudaStream_t st = ctx->stream;
  const bool ts = !(std::isinf(eta) && eta > 0);
  if (!ctx->lft_attr_set) {
    CK(cudaFuncSetAttribute(lf_tree_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LFT_SMEM));
    CK(cudaFuncSetAttribute(lf_tree_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LFT_SMEM));
    CK(cudaFuncSetAttribute(lf_tree_kernel<true, LFT_CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)LFT_SMEM));
    CK(cudaFuncSetAttribute(lf_tree_kernel<false, LFT_CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)LFT_SMEM));
    ctx->lft_attr_set = true;
  }
  CK(cudaMemcpyAsync(w, t, (size_t)iend * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(w + LFT_MAX, 0, 2 * sizeof(double), st));
  LfTreeParams P;
  P.packed = (double *)ctx->d_lf_packed.p;
  P.dt_max = (double *)ctx->d_lf_dtmax.p;
  P.t = w;
  P.out = (long long *)(w + LFT_MAX);
  P.map = (int *)(w + LFT_MAX + 2);
  P.iend = iend;
  P.dt = dt;
  P.eta = eta;
  // narrow prefixes: one CTA; wider ones: a cluster of LFT_CLUSTER CTAs that share out the targets
  // of every kick block (NBK_LFT_CLUSTER=0/1 forces either - a knob for tests and profiles)
  bool use_cluster = iend > 96;
  if (const char *e = getenv("NBK_LFT_CLUSTER")) use_cluster = atoi(e) != 0;
  if (use_cluster) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(LFT_CLUSTER);
    cfg.blockDim = dim3(LFT_THREADS);
    cfg.dynamicSmemBytes = LFT_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = LFT_CLUSTER;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    void *args[] = {&P};
    const void *kern = ts ? (const void *)lf_tree_kernel<true, LFT_CLUSTER>
                          : (const void *)lf_tree_kernel<false, LFT_CLUSTER>;
    CK(cudaLaunchKernelExC(&cfg, kern, args));
  } else if (ts) {
    lf_tree_kernel<true, 1><<<1, LFT_THREADS, LFT_SMEM, st>>>(P);
  } else {
    lf_tree_kernel<false, 1><<<1, LFT_THREADS, LFT_SMEM, st>>>(P);
  }
  CK(cudaGetLastError());
  ctx->launches++;
  long long out[2] = {0, 0};
  CK(cudaMemcpyAsync(t, P.t, (size_t)iend * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(dt_max, P.dt_max, (size_t)iend * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(perm, P.map, (size_t)iend * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(out, P.out, sizeof out, cudaMemcpyDeviceToHost, st));
  TRY(sync(ctx));
  if (out[1] != 0)
    return fail(ctx, NBK_ERR_STATE, "nbk_lf_advance_subsystem: block recursion deeper than %d frames",
                LFT_DEPTH);
  if (nkicks) *nkicks += out[0];
  return NBK_OK;
}

// ---- device-resident Omelyan_advancer.OA ---------------------------------------------------------
int nbk_oa_upload(nbk_ctx *ctx, int n, const double *m, const double *q, const double *p) {
  if (!ctx) return NBK_ERR_ARG;
  if (n < 0 || (n && (!m || !q || !p))) return fail(ctx, NBK_ERR_ARG, "nbk_oa_upload: bad arguments");
  CK(cudaSetDevice(ctx->device));
  TRY(upload(ctx, ctx->d_oa_m, m, n));
  TRY(upload(ctx, ctx->d_oa_q, q, 3 * (size_t)n));
  TRY(upload(ctx, ctx->d_oa_p, p, 3 * (size_t)n));
  TRY(reserve(ctx, ctx->d_oa_qs, 3 * (size_t)n * sizeof(double)));
  TRY(reserve(ctx, ctx->d_oa_g, 3 * (size_t)n * sizeof(double)));
  TRY(reserve(ctx, ctx->d_packed, (size_t)n * PK * sizeof(double)));
  ctx->oa_n = n;
  ctx->oa_g_valid = false;
  return sync(ctx);
}

namespace {
// gsum of the resident bodies at positions `pos` into d_oa_g
int oa_sweep(nbk_ctx *ctx, const double *pos, double eps2) {
  const int n = ctx->oa_n;
  TRY(pack(ctx, n, (const double *)ctx->d_oa_m.p, pos, nullptr, 1, (double *)ctx->d_packed.p,
           ctx->stream));
  SweepParams P = base_params((const double *)ctx->d_packed.p, eps2, 0, n, 0, n);
  P.out[0] = (double *)ctx->d_oa_g.p;
  return K1::run(ctx, P, ctx->stream, true);
}
}  // namespace

int nbk_oa_steps(nbk_ctx *ctx, double h, double eps2, int nsteps) {
  if (!ctx) return NBK_ERR_ARG;
  if (ctx->oa_n <= 0 || nsteps < 0) return fail(ctx, NBK_ERR_STATE, "nbk_oa_steps: no resident bodies");
  CK(cudaSetDevice(ctx->device));
  const int n = ctx->oa_n;
  const size_t n3 = 3 * (size_t)n;
  double *m = (double *)ctx->d_oa_m.p, *q = (double *)ctx->d_oa_q.p, *p = (double *)ctx->d_oa_p.p;
  double *qs = (double *)ctx->d_oa_qs.p, *g = (double *)ctx->d_oa_g.p;
  const unsigned b3 = (unsigned)((n3 + 255) / 256), b1 = (unsigned)((n + 255) / 256);
  cudaStream_t st = ctx->stream;
  for (int s = 0; s < nsteps; ++s) {
    if (n < 2) {  // no pairs: only the two a_operate drifts act
      oa_drift_kernel<<<b1, 256, 0, st>>>(n, m, q, p, h);
      oa_drift_kernel<<<b1, 256, 0, st>>>(n, m, q, p, h);
      ctx->launches += 2;
      continue;
    }
    // B(h): the step's first b_operate sees the positions the previous step's last one saw
    // (nothing moves q in between, omelyan_advancer.ml:96-104), so its sums are reused
    if (!(ctx->oa_g_valid && ctx->oa_g_eps2 == eps2)) TRY(oa_sweep(ctx, q, eps2));
    oa_axpy_kernel<<<b3, 256, 0, st>>>(p, g, n3, h / 6.0);
    oa_drift_kernel<<<b1, 256, 0, st>>>(n, m, q, p, h);
    // C(h)
    TRY(oa_sweep(ctx, q, eps2));
    oa_qstar_kernel<<<b1, 256, 0, st>>>(n, m, q, g, h, qs);
    TRY(oa_sweep(ctx, qs, eps2));
    oa_axpy_kernel<<<b3, 256, 0, st>>>(p, g, n3, 2.0 * h / 3.0);
    oa_drift_kernel<<<b1, 256, 0, st>>>(n, m, q, p, h);
    // B(h)
    TRY(oa_sweep(ctx, q, eps2));
    oa_axpy_kernel<<<b3, 256, 0, st>>>(p, g, n3, h / 6.0);
    CK(cudaGetLastError());
    ctx->launches += 6;
    ctx->oa_g_valid = true;
    ctx->oa_g_eps2 = eps2;
  }
  return NBK_OK;
}

int nbk_oa_download(nbk_ctx *ctx, double *q, double *p) {
  if (!ctx) return NBK_ERR_ARG;
  if (ctx->oa_n <= 0) return fail(ctx, NBK_ERR_STATE, "nbk_oa_download: no resident bodies");
  CK(cudaSetDevice(ctx->device));
  if (q) TRY(download(ctx, q, ctx->d_oa_q, 3 * (size_t)ctx->oa_n));
  if (p) TRY(download(ctx, p, ctx->d_oa_p, 3 * (size_t)ctx->oa_n));
  return sync(ctx);
}

int nbk_oa_diagnostics(nbk_ctx *ctx, double eps2, double *ke, double *pe, double *el) {
  if (!ctx) return NBK_ERR_ARG;
  if (ctx->oa_n <= 0) return fail(ctx, NBK_ERR_STATE, "nbk_oa_diagnostics: no resident bodies");
  CK(cudaSetDevice(ctx->device));
  const int n = ctx->oa_n;
  double *m = (double *)ctx->d_oa_m.p, *q = (double *)ctx->d_oa_q.p, *p = (double *)ctx->d_oa_p.p;
  TRY(reserve(ctx, ctx->d_oa_phi, (size_t)n * sizeof(double)));
  TRY(reserve(ctx, ctx->d_oa_el, 4 * (size_t)n * sizeof(double)));
  TRY(reserve(ctx, ctx->d_out[0], (size_t)n * sizeof(double)));
  TRY(reserve(ctx, ctx->d_scalar, 64));
  double *phi = (double *)ctx->d_oa_phi.p, *sc = (double *)ctx->d_scalar.p;
  cudaStream_t st = ctx->stream;
  // one potential sweep serves Energy.total_potential_energy (0.5 sum phi) and every E_i
  TRY(pack(ctx, n, m, q, nullptr, 1, (double *)ctx->d_packed.p, st));
  if (n >= 2) {
    SweepParams P = base_params((const double *)ctx->d_packed.p, eps2, 0, n, 0, n);
    P.out[0] = phi;
    TRY(K3::run(ctx, P, st, true));
  } else {
    CK(cudaMemsetAsync(phi, 0, (size_t)n * sizeof(double), st));
  }
  sum_kernel<<<1, 1024, 0, st>>>(phi, n, sc);
  kinetic_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, m, p, (double *)ctx->d_out[0].p);
  sum_kernel<<<1, 1024, 0, st>>>((const double *)ctx->d_out[0].p, n, sc + 1);
  ctx->launches += 3;
  if (el) {
    el_phase_space_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, m, q, p, phi, (double *)ctx->d_oa_el.p);
    ctx->launches++;
    TRY(download(ctx, el, ctx->d_oa_el, 4 * (size_t)n));
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->h_scalar, sc, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TRY(sync(ctx));
  if (pe) *pe = 0.5 * ctx->h_scalar[0];
  if (ke) *ke = ctx->h_scalar[1];
  return NBK_OK;
}

// ---- device-resident Hermite state -------------------------------------------------------------
int nbk_hermite_upload(nbk_ctx *ctx, int n, const double *t, const double *m, const double *q,
                       const double *p, const double *hh, const double *f, const double *df) {
  if (!ctx) return NBK_ERR_ARG;
  if (n <= 0 || !t || !m || !q || !p || !f || !df)
    return fail(ctx, NBK_ERR_ARG, "nbk_hermite_upload: bad arguments");
  CK(cudaSetDevice(ctx->device));
  std::string buf((size_t)n * HB * sizeof(double), '\0');
  double *h = reinterpret_cast<double *>(&buf[0]);
  for (int i = 0; i < n; ++i) {
    double *r = h + (size_t)i * HB;
    r[0] = t[i];
    r[1] = m[i];
    for (int k = 0; k < 3; ++k) {
      r[2 + k] = q[3 * i + k];
      r[5 + k] = p[3 * i + k];
      r[8 + k] = f[3 * i + k];
      r[11 + k] = df[3 * i + k];
    }
    r[14] = hh ? hh[i] : 0.0;
    r[15] = 0.0;  // re-insertion stamp (hermite_resident_kernel)
  }
  const int nblk = (n + 255) / 256;
  TRY(reserve(ctx, ctx->d_hstate, (size_t)n * HB * sizeof(double)));
  TRY(reserve(ctx, ctx->d_hpartial, (size_t)nblk * 6 * sizeof(double)));
  TRY(reserve(ctx, ctx->d_hticket, 64));
  TRY(reserve(ctx, ctx->d_hout, 64));
  CK(cudaMemcpyAsync(ctx->d_hstate.p, h, (size_t)n * HB * sizeof(double), cudaMemcpyHostToDevice,
                     ctx->stream));
  CK(cudaMemsetAsync(ctx->d_hticket.p, 0, 64, ctx->stream));
  TRY(sync(ctx));
  ctx->hermite_n = n;
  return NBK_OK;
}

int nbk_hermite_body_sum(nbk_ctx *ctx, int i, double nt, double eps2, int upd,
                         const double *upd_state, double *gsum, double *dgsum) {
  if (!ctx) return NBK_ERR_ARG;
  const int n = ctx->hermite_n;
  if (n == 0) return fail(ctx, NBK_ERR_STATE, "no resident Hermite state: call nbk_hermite_upload");
  if (i < 0 || i >= n || upd >= n || (upd >= 0 && !upd_state) || !gsum || !dgsum)
    return fail(ctx, NBK_ERR_ARG, "nbk_hermite_body_sum: bad arguments (i=%d upd=%d n=%d)", i, upd, n);
  CK(cudaSetDevice(ctx->device));
  HermiteUpd u;
  memset(&u, 0, sizeof u);
  if (upd >= 0) memcpy(u.s, upd_state, sizeof u.s);
  const int nblk = (n + 255) / 256;
  if (!ctx->h_hres) {
    CK(cudaHostAlloc((void **)&ctx->h_hres, 64, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void **)&ctx->d_hres, ctx->h_hres, 0));
    memset(ctx->h_hres, 0, 64);
  }
  const double seq = (ctx->hseq += 1.0);
  hermite_body_kernel<<<nblk, 256, 0, ctx->stream>>>(
      (double *)ctx->d_hstate.p, n, i, nt, eps2, upd, u, (double *)ctx->d_hpartial.p,
      (unsigned int *)ctx->d_hticket.p, ctx->d_hres, seq);
  CK(cudaGetLastError());
  ctx->launches++;
  // spin on the sequence number the kernel writes last; fall back to the stream on errors
  volatile double *flag = ctx->h_hres + 7;
  for (unsigned long spins = 0; *flag != seq; ++spins) {
    if ((spins & 0xfffff) == 0xfffff) {
      cudaError_t q = cudaStreamQuery(ctx->stream);
      if (q != cudaSuccess && q != cudaErrorNotReady)
        return fail(ctx, NBK_ERR_CUDA, "hermite_body_kernel failed: %s", cudaGetErrorString(q));
      if (q == cudaSuccess && *flag != seq)
        return fail(ctx, NBK_ERR_CUDA, "hermite_body_kernel finished without publishing");
    }
  }
  for (int k = 0; k < 3; ++k) {
    gsum[k] = ctx->h_hres[k];
    dgsum[k] = ctx->h_hres[3 + k];
  }
  return NBK_OK;
}

int nbk_hermite_resident_max(void) { return HGR_MAX_N; }

int nbk_hermite_advance_resident(nbk_ctx *ctx, double stop_time, double eta, double eps2,
                                 long max_steps, long *nsteps) {
  if (!ctx) return NBK_ERR_ARG;
  const int n = ctx->hermite_n;
  if (n == 0) return fail(ctx, NBK_ERR_STATE, "no resident Hermite state: call nbk_hermite_upload");
  if (n > HGR_MAX_N || n > ctx->sm_count * HGR_PER_CTA)
    return fail(ctx, NBK_ERR_ARG, "nbk_hermite_advance_resident: n=%d exceeds %d", n,
                std::min(HGR_MAX_N, ctx->sm_count * HGR_PER_CTA));
  CK(cudaSetDevice(ctx->device));
  TRY(reserve(ctx, ctx->d_scalar, 64));
  if (n > HCL_MAX_N || ctx->hermite_threads == -148) {
    // the whole grid, one CTA per SM, state split over their shared memories, candidates and
    // partial sums through L2, two grid barriers per step (cooperative launch)
    int G = std::min(ctx->sm_count, HGR_MAX_CTAS);
    G = std::min(G, std::max((n + HGR_THREADS - 1) / HGR_THREADS, (n + HGR_PER_CTA - 1) / HGR_PER_CTA));
    if (const char *e = getenv("NBK_HERMITE_GRID_CTAS")) G = std::max(1, std::min(G, atoi(e)));
    G = std::max(G, (n + HGR_PER_CTA - 1) / HGR_PER_CTA);
    const int per = (n + G - 1) / G;
    const size_t gsmem = (size_t)per * HRES_STRIDE * sizeof(double);
    if (ctx->hermite_spec) {  // software-pipelined (default); NBK_HERMITE_SPEC=0: the plain loop
      const size_t ncand = (size_t)2 * G + 1;
      const size_t swords = 2 * ncand * HCL_CAND + (size_t)2 * G * 8 + 8;
      TRY(reserve(ctx, ctx->d_hgrid, swords * sizeof(double)));
      double *sw = (double *)ctx->d_hgrid.p;
      HermiteGridSpecParams S;
      S.state = (double *)ctx->d_hstate.p;
      S.cand = sw;
      S.part = sw + 2 * ncand * HCL_CAND;
      S.bar = (unsigned long long *)(S.part + (size_t)2 * G * 8);
      S.nsteps_out = (long long *)ctx->d_scalar.p;
      S.n = n;
      S.stop_time = stop_time;
      S.eta = eta;
      S.eps2 = eps2;
      S.max_steps = (long long)max_steps;
      CK(cudaMemsetAsync(S.bar, 0, 8 * sizeof(double), ctx->stream));
      CK(cudaFuncSetAttribute(hermite_grid_spec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)gsmem));
      void *sargs[] = {&S};
      CK(cudaLaunchCooperativeKernel((const void *)hermite_grid_spec_kernel, dim3(G),
                                     dim3(HGRS_THREADS), sargs, gsmem, ctx->stream));
      ctx->launches++;
      CK(cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar.p, sizeof(long long), cudaMemcpyDeviceToHost,
                         ctx->stream));
      TRY(sync(ctx));
      long long ssteps;
      memcpy(&ssteps, ctx->h_scalar, sizeof ssteps);
      if (nsteps) *nsteps = (long)ssteps;
      if (ssteps >= (long long)max_steps && max_steps > 0)
        return fail(ctx, NBK_ERR_STATE, "Hermite step cap (%ld) reached before stop_time", max_steps);
      return NBK_OK;
    }
    const size_t words = (size_t)2 * G * HCL_CAND + (size_t)2 * G * 8 + 8;
    TRY(reserve(ctx, ctx->d_hgrid, words * sizeof(double)));
    double *w = (double *)ctx->d_hgrid.p;
    HermiteGridParams Q;
    Q.state = (double *)ctx->d_hstate.p;
    Q.cand = w;
    Q.part = w + (size_t)2 * G * HCL_CAND;
    Q.bar = (unsigned long long *)(Q.part + (size_t)2 * G * 8);
    Q.nsteps_out = (long long *)ctx->d_scalar.p;
    Q.n = n;
    Q.stop_time = stop_time;
    Q.eta = eta;
    Q.eps2 = eps2;
    Q.max_steps = (long long)max_steps;
    CK(cudaMemsetAsync(Q.bar, 0, 8 * sizeof(double), ctx->stream));
    CK(cudaFuncSetAttribute(hermite_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)gsmem));
    void *args[] = {&Q};
    CK(cudaLaunchCooperativeKernel((const void *)hermite_grid_kernel, dim3(G), dim3(HGR_THREADS),
                                   args, gsmem, ctx->stream));
    ctx->launches++;
    CK(cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar.p, sizeof(long long), cudaMemcpyDeviceToHost,
                       ctx->stream));
    TRY(sync(ctx));
    long long gsteps;
    memcpy(&gsteps, ctx->h_scalar, sizeof gsteps);
    if (nsteps) *nsteps = (long)gsteps;
    if (gsteps >= (long long)max_steps && max_steps > 0)
      return fail(ctx, NBK_ERR_STATE, "Hermite step cap (%ld) reached before stop_time", max_steps);
    return NBK_OK;
  }
  if (n > HRES_MAX_N || ctx->hermite_threads == -8) {
    // thread-block cluster of 8 CTAs, state split over their shared memories (DSMEM exchange)
    const int per = (n + HCL_CTAS - 1) / HCL_CTAS;
    const size_t csmem = (size_t)per * HRES_STRIDE * sizeof(double);
    CK(cudaFuncSetAttribute(hermite_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)csmem));
    if (ctx->hermite_spec) {  // software-pipelined (default); NBK_HERMITE_SPEC=0: the plain loop
      CK(cudaFuncSetAttribute(hermite_cluster_spec_kernel,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
      static const bool want_prof = getenv("NBK_HERMITE_PROF") != nullptr;
      long long *d_prof = nullptr;
      if (want_prof) {
        TRY(reserve(ctx, ctx->d_hgrid, 8 * sizeof(long long)));
        d_prof = (long long *)ctx->d_hgrid.p;
        CK(cudaMemsetAsync(d_prof, 0, 8 * sizeof(long long), ctx->stream));
      }
      hermite_cluster_spec_kernel<<<HCL_CTAS, HCLS_THREADS, csmem, ctx->stream>>>(
          (double *)ctx->d_hstate.p, n, stop_time, eta, eps2, (long long)max_steps,
          (long long *)ctx->d_scalar.p, d_prof);
      if (want_prof) {
        long long hp[8];
        CK(cudaMemcpyAsync(hp, d_prof, sizeof hp, cudaMemcpyDeviceToHost, ctx->stream));
        TRY(sync(ctx));
        fprintf(stderr, "hermite cluster prof, CTA 0 (cycles): top %lld | corrector %lld (%lld steps owned) | "
                        "search+post %lld | sums+post %lld | barrier %lld | pick %lld\n",
                hp[0], hp[1], hp[6], hp[2], hp[3], hp[4], hp[5]);
      }
    } else {
      hermite_cluster_kernel<<<HCL_CTAS, HCL_THREADS, csmem, ctx->stream>>>(
          (double *)ctx->d_hstate.p, n, stop_time, eta, eps2, (long long)max_steps,
          (long long *)ctx->d_scalar.p);
    }
    CK(cudaGetLastError());
    ctx->launches++;
    CK(cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar.p, sizeof(long long), cudaMemcpyDeviceToHost,
                       ctx->stream));
    TRY(sync(ctx));
    long long csteps;
    memcpy(&csteps, ctx->h_scalar, sizeof csteps);
    if (nsteps) *nsteps = (long)csteps;
    if (csteps >= (long long)max_steps && max_steps > 0)
      return fail(ctx, NBK_ERR_STATE, "Hermite step cap (%ld) reached before stop_time", max_steps);
    return NBK_OK;
  }
  const size_t smem = (size_t)n * HRES_STRIDE * sizeof(double);
  // Threads of the single stepping CTA: the per-step cost is dominated by barriers and the two
  // block reductions, so fewer, busier warps win (measured on B200: profiles/).
  // the pipelined loop wants ten warps at N = 1024 (two summing warps per scheduler): measured
  // 2.02 us per step with 320 threads, 2.12 with 256, 2.15 with 384; N = 512: 1.70 with 192;
  // N = 256: 1.45 - 1.49 with 128 or 256; N = 1536: 2.44 with 384
  int nth = ctx->hermite_threads;
  if (nth <= 0)
    nth = ctx->hermite_spec && n >= 64 ? (n <= 256 ? 128 : n <= 640 ? 192 : n <= 1152 ? 320 : 384)
                                      : (n <= 128 ? 128 : 256);
  auto launch = [&](auto kern, int threads) -> int {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<1, threads, smem, ctx->stream>>>((double *)ctx->d_hstate.p, n, stop_time, eta, eps2,
                                            (long long)max_steps, (long long *)ctx->d_scalar.p);
    return NBK_OK;
  };
  // NBK_HERMITE_SPEC=0: the plain loop; default: the software-pipelined one (corrector of step k
  // overlapped with the search and the force sum of step k+1)
  const bool spec = ctx->hermite_spec;
  if (spec && n >= 64) {
    static const bool want_prof = getenv("NBK_HERMITE_PROF") != nullptr;
    long long *d_prof = nullptr;
    if (want_prof) {
      TRY(reserve(ctx, ctx->d_hgrid, 8 * sizeof(long long)));
      d_prof = (long long *)ctx->d_hgrid.p;
      CK(cudaMemsetAsync(d_prof, 0, 8 * sizeof(long long), ctx->stream));
    }
    auto launch_spec = [&](auto kern, int threads) -> int {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<1, threads, smem, ctx->stream>>>((double *)ctx->d_hstate.p, n, stop_time, eta, eps2,
                                              (long long)max_steps, (long long *)ctx->d_scalar.p,
                                              d_prof);
      return NBK_OK;
    };
    switch (nth) {
      case 128: TRY(launch_spec(hermite_resident_spec_kernel<128>, 128)); break;
      case 192: TRY(launch_spec(hermite_resident_spec_kernel<192>, 192)); break;
      case 320: TRY(launch_spec(hermite_resident_spec_kernel<320>, 320)); break;
      case 384: TRY(launch_spec(hermite_resident_spec_kernel<384>, 384)); break;
      case 512: TRY(launch_spec(hermite_resident_spec_kernel<512>, 512)); break;
      case 1024: TRY(launch_spec(hermite_resident_spec_kernel<1024>, 1024)); break;
      default: TRY(launch_spec(hermite_resident_spec_kernel<256>, 256)); break;
    }
    if (want_prof) {
      long long hp[8];
      CK(cudaMemcpyAsync(hp, d_prof, sizeof hp, cudaMemcpyDeviceToHost, ctx->stream));
      TRY(sync(ctx));
      fprintf(stderr, "hermite spec prof (cycles): t0 plain-sum %lld corrector %lld barrier %lld | "
                      "t32 select %lld spec-sum %lld barrier %lld | plain steps %lld\n",
              hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[6]);
    }
  } else {
    switch (nth) {
      case 128: TRY(launch(hermite_resident_kernel<128>, 128)); break;
      case 512: TRY(launch(hermite_resident_kernel<512>, 512)); break;
      case 1024: TRY(launch(hermite_resident_kernel<1024>, 1024)); break;
      default: TRY(launch(hermite_resident_kernel<256>, 256)); break;
    }
  }
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar.p, sizeof(long long), cudaMemcpyDeviceToHost,
                     ctx->stream));
  TRY(sync(ctx));
  long long steps;
  memcpy(&steps, ctx->h_scalar, sizeof steps);
  if (nsteps) *nsteps = (long)steps;
  if (steps >= (long long)max_steps && max_steps > 0)
    return fail(ctx, NBK_ERR_STATE, "Hermite step cap (%ld) reached before stop_time", max_steps);
  return NBK_OK;
}

int nbk_hermite_download(nbk_ctx *ctx, double *t, double *q, double *p, double *h, double *f,
                         double *df) {
  if (!ctx) return NBK_ERR_ARG;
  const int n = ctx->hermite_n;
  if (n == 0) return fail(ctx, NBK_ERR_STATE, "no resident Hermite state");
  CK(cudaSetDevice(ctx->device));
  std::string buf((size_t)n * HB * sizeof(double), '\0');
  double *hb = reinterpret_cast<double *>(&buf[0]);
  CK(cudaMemcpyAsync(hb, ctx->d_hstate.p, (size_t)n * HB * sizeof(double), cudaMemcpyDeviceToHost,
                     ctx->stream));
  TRY(sync(ctx));
  for (int i = 0; i < n; ++i) {
    const double *r = hb + (size_t)i * HB;
    if (t) t[i] = r[0];
    if (h) h[i] = r[14];
    for (int k = 0; k < 3; ++k) {
      if (q) q[3 * i + k] = r[2 + k];
      if (p) p[3 * i + k] = r[5 + k];
      if (f) f[3 * i + k] = r[8 + k];
      if (df) df[3 * i + k] = r[11 + k];
    }
  }
  return NBK_OK;
}

// ---- raw device-pointer API ----------------------------------------------------------------------
int nbk_dev_pack(nbk_ctx *ctx, int n, const double *d_m, const double *d_q, const double *d_vel,
                 int is_velocity, double *d_packed, void *stream) {
  if (!ctx) return NBK_ERR_ARG;
  if (n < 0 || !d_m || !d_q || !d_packed || ((uintptr_t)d_packed & 15))
    return fail(ctx, NBK_ERR_ARG, "nbk_dev_pack: bad arguments (d_packed must be 16-byte aligned)");
  CK(cudaSetDevice(ctx->device));
  return pack(ctx, n, d_m, d_q, d_vel, is_velocity, d_packed, (cudaStream_t)stream);
}

int nbk_dev_pack_push(nbk_ctx *ctx, int n, const double *d_m, const double *d_q,
                      const double *d_vel, int is_velocity, int nranks, double *const *d_dst,
                      void *stream) {
  if (!ctx) return NBK_ERR_ARG;
  if (n < 0 || !d_m || !d_q || !d_dst || nranks < 1 || nranks > MAX_PEERS)
    return fail(ctx, NBK_ERR_ARG, "nbk_dev_pack_push: bad arguments");
  PackPushParams D;
  memset(&D, 0, sizeof D);
  for (int r = 0; r < nranks; ++r) {
    if (!d_dst[r] || ((uintptr_t)d_dst[r] & 15))
      return fail(ctx, NBK_ERR_ARG, "nbk_dev_pack_push: destination %d null or misaligned", r);
    D.dst[r] = d_dst[r];
  }
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return NBK_OK;
  pack_push_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, d_m, d_q, d_vel,
                                                                      is_velocity, nranks, D);
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;
}

#define DEV_COMMON()                                                                     \
  do {                                                                                   \
    TRY(check_ranges(ctx, n_total, t0, t1, s0, s1));                                     \
    if (!d_packed || ((uintptr_t)d_packed & 15))                                         \
      return fail(ctx, NBK_ERR_ARG, "d_packed must be a 16-byte aligned device pointer"); \
    CK(cudaSetDevice(ctx->device));                                                      \
    if (t1 == t0 || s1 == s0) return NBK_OK;                                             \
  } while (0)

int nbk_dev_grad_v_dgrad_v_sum(nbk_ctx *ctx, const double *d_packed, int n_total, double eps2,
                               int t0, int t1, int s0, int s1, double *d_gsum, double *d_dgsum,
                               int accumulate, void *stream) {
  DEV_COMMON();
  if (!d_gsum || !d_dgsum) return fail(ctx, NBK_ERR_ARG, "null output");
  SweepParams P = base_params(d_packed, eps2, t0, t1, s0, s1);
  P.out[0] = d_gsum;
  P.out[1] = d_dgsum;
  P.accumulate = accumulate;
  cudaStream_t st = (cudaStream_t)stream;
  return run_k2(ctx, P, st, false);
}

int nbk_set_sym(nbk_ctx *ctx, int mode) {
  if (!ctx) return NBK_ERR_ARG;
  if (mode < 0 || mode > 2) return fail(ctx, NBK_ERR_ARG, "nbk_set_sym: mode must be 0, 1 or 2");
  ctx->sym_mode = mode;
  for (nbk_ctx *p : ctx->peers) p->sym_mode = mode;
  return NBK_OK;
}

int nbk_sym_items(int n_total, long long *items) {
  if (n_total < 0 || !items) return NBK_ERR_ARG;
  *items = sym_items(n_total);
  return NBK_OK;
}

int nbk_dev_grad_v_dgrad_v_sym(nbk_ctx *ctx, const double *d_packed, int n_total, double eps2,
                               long long item0, long long item1, double *d_sum6, void *stream) {
  if (!ctx) return NBK_ERR_ARG;
  if (n_total < 0 || !d_packed || ((uintptr_t)d_packed & 15) || !d_sum6 || item0 < 0 ||
      item1 < item0 || item1 > sym_items(n_total))
    return fail(ctx, NBK_ERR_ARG, "nbk_dev_grad_v_dgrad_v_sym: bad arguments");
  CK(cudaSetDevice(ctx->device));
  if (n_total == 0) return NBK_OK;
  const long long rows_cap = (((long long)n_total + SYM_B - 1) / SYM_B) * SYM_B;
  return sym_run(ctx, &d_packed, rows_cap, 1, 0, nullptr, 0, n_total, eps2, item0, item1, d_sum6,
                 nullptr, nullptr, 0, n_total, (cudaStream_t)stream, false);
}

int nbk_dev_sym_gather_peer(nbk_ctx *ctx, int nranks, int self, const double *const *d_sum6_of_rank,
                            const uint64_t *d_ready, uint64_t epoch, int b0, int b1,
                            double *d_gsum, double *d_dgsum, void *stream) {
  if (!ctx) return NBK_ERR_ARG;
  if (nranks < 1 || nranks > MAX_PEERS || self < 0 || self >= nranks || !d_sum6_of_rank ||
      b0 < 0 || b1 < b0 || !d_gsum || !d_dgsum)
    return fail(ctx, NBK_ERR_ARG, "nbk_dev_sym_gather_peer: bad arguments");
  SymGatherParams G;
  memset(&G, 0, sizeof G);
  for (int r = 0; r < nranks; ++r) {
    if (!d_sum6_of_rank[r]) return fail(ctx, NBK_ERR_ARG, "nbk_dev_sym_gather_peer: null sums of rank %d", r);
    G.sum6[r] = d_sum6_of_rank[r];
  }
  G.ndev = nranks;
  G.self = self;
  G.a = b0;
  G.b = b1;
  G.g = d_gsum;
  G.dg = d_dgsum;
  G.ready = (const unsigned long long *)d_ready;
  G.epoch = epoch;
  G.wait_ns = ctx->peer_wait_ns;
  CK(cudaSetDevice(ctx->device));
  if (b1 == b0) return NBK_OK;
  sym_gather_kernel<<<((b1 - b0) * 6 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(G);
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;
}

int nbk_dev_split6(nbk_ctx *ctx, const double *d_sum6, int count, double *d_gsum, double *d_dgsum,
                   void *stream) {
  if (!ctx) return NBK_ERR_ARG;
  if (count < 0 || !d_sum6 || !d_gsum || !d_dgsum)
    return fail(ctx, NBK_ERR_ARG, "nbk_dev_split6: bad arguments");
  CK(cudaSetDevice(ctx->device));
  if (count == 0) return NBK_OK;
  sym_split_kernel<<<(count * 6 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_sum6, count, d_gsum,
                                                                           d_dgsum);
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;
}

int nbk_dev_grad_v_dgrad_v_sym_peer(nbk_ctx *ctx, const nbk_shards *sh, int n_total, double eps2,
                                    long long item0, long long item1, double *d_sum6,
                                    void *stream) {
  if (!ctx) return NBK_ERR_ARG;
  if (!sh || sh->nranks < 1 || sh->nranks > MAX_PEERS || sh->self < 0 || sh->self >= sh->nranks ||
      sh->shard_rows <= 0 || sh->shard_rows % SYM_B != 0 ||
      (long long)sh->shard_rows * sh->nranks < n_total)
    return fail(ctx, NBK_ERR_ARG,
                "nbk_dev_grad_v_dgrad_v_sym_peer: bad shard table (shard_rows a multiple of %d)",
                SYM_B);
  if (n_total < 0 || !d_sum6 || item0 < 0 || item1 < item0 || item1 > sym_items(n_total))
    return fail(ctx, NBK_ERR_ARG, "nbk_dev_grad_v_dgrad_v_sym_peer: bad arguments");
  for (int r = 0; r < sh->nranks; ++r)
    if (!sh->packed[r] || ((uintptr_t)sh->packed[r] & 15))
      return fail(ctx, NBK_ERR_ARG, "shard %d: null or misaligned device pointer", r);
  CK(cudaSetDevice(ctx->device));
  if (n_total == 0) return NBK_OK;
  return sym_run(ctx, sh->packed, sh->shard_rows, sh->nranks, sh->self,
                 (const unsigned long long *)sh->ready, sh->epoch, n_total, eps2, item0, item1,
                 d_sum6, nullptr, nullptr, 0, n_total, (cudaStream_t)stream, false);
}

int nbk_dev_grad_v_sum(nbk_ctx *ctx, const double *d_packed, int n_total, double eps2, int t0,
                       int t1, int s0, int s1, double *d_gsum, int accumulate, void *stream) {
  DEV_COMMON();
  if (!d_gsum) return fail(ctx, NBK_ERR_ARG, "null output");
  SweepParams P = base_params(d_packed, eps2, t0, t1, s0, s1);
  P.out[0] = d_gsum;
  P.accumulate = accumulate;
  cudaStream_t st = (cudaStream_t)stream;
  return K1::run(ctx, P, st, false);
}

int nbk_dev_v_sum(nbk_ctx *ctx, const double *d_packed, int n_total, double eps2, int t0, int t1,
                  int s0, int s1, double *d_phi, int accumulate, void *stream) {
  DEV_COMMON();
  if (!d_phi) return fail(ctx, NBK_ERR_ARG, "null output");
  SweepParams P = base_params(d_packed, eps2, t0, t1, s0, s1);
  P.out[0] = d_phi;
  P.accumulate = accumulate;
  cudaStream_t st = (cudaStream_t)stream;
  return K3::run(ctx, P, st, false);
}

int nbk_dev_kick_sum(nbk_ctx *ctx, const double *d_packed, int n_total, double eta, double dt,
                     int t0, int t1, int s0, int s1, double *d_dv, double *d_tmin, void *stream) {
  DEV_COMMON();
  if (!d_dv) return fail(ctx, NBK_ERR_ARG, "null output");
  SweepParams P = base_params(d_packed, 0.0, t0, t1, s0, s1);
  P.dt = dt;
  P.eta = eta;
  P.out[0] = d_dv;
  P.out[1] = d_tmin;
  cudaStream_t st = (cudaStream_t)stream;
  return d_tmin ? K4T::run(ctx, P, st, false) : K4::run(ctx, P, st, false);
}

int nbk_dev_pack_aux(nbk_ctx *ctx, int n, const double *d_m, const double *d_dq,
                     const double *d_dp, double *d_aux, void *stream) {
  if (!ctx) return NBK_ERR_ARG;
  if (n < 0 || !d_m || !d_dq || !d_aux || ((uintptr_t)d_aux & 15))
    return fail(ctx, NBK_ERR_ARG, "nbk_dev_pack_aux: bad arguments");
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return NBK_OK;
  pack_aux_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, d_m, d_dq, d_dp, 0, d_aux);
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;
}

int nbk_dev_predict_pack(nbk_ctx *ctx, int n, const double *d_t, const double *d_m,
                         const double *d_q, const double *d_p, const double *d_f,
                         const double *d_df, double nt, double *d_packed, void *stream) {
  if (!ctx) return NBK_ERR_ARG;
  if (n < 0 || (n > 0 && (!d_t || !d_m || !d_q || !d_p || !d_f || !d_df || !d_packed)))
    return fail(ctx, NBK_ERR_ARG, "nbk_dev_predict_pack: bad arguments");
  if (n == 0) return NBK_OK;
  CK(cudaSetDevice(ctx->device));
  predict_pack_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, d_t, d_m, d_q, d_p, d_f,
                                                                        d_df, nt, d_packed);
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;
}

int nbk_dev_predict_pack_tangent(nbk_ctx *ctx, int n, int K, const double *d_t, const double *d_m,
                                 const double *d_q, const double *d_p, const double *d_f,
                                 const double *d_df, double nt, const double *d_t_tan,
                                 const double *d_q_tan, const double *d_p_tan,
                                 const double *d_f_tan, const double *d_df_tan,
                                 const double *d_nt_tan, double *d_aux, size_t aux_stride,
                                 void *stream) {
  if (!ctx) return NBK_ERR_ARG;
  if (n < 0 || K < 0 ||
      (n > 0 && K > 0 &&
       (!d_t || !d_m || !d_q || !d_p || !d_f || !d_df || !d_q_tan || !d_p_tan || !d_f_tan ||
        !d_df_tan || !d_aux || aux_stride < (size_t)n * PK)))
    return fail(ctx, NBK_ERR_ARG, "nbk_dev_predict_pack_tangent: bad arguments");
  if (n == 0 || K == 0) return NBK_OK;
  CK(cudaSetDevice(ctx->device));
  dim3 grid((n + 255) / 256, K);
  predict_pack_tangent_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      n, d_t, d_m, d_q, d_p, d_f, d_df, nt, d_t_tan, d_q_tan, d_p_tan, d_f_tan, d_df_tan, d_nt_tan,
      d_aux, aux_stride);
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;
}

int nbk_dev_grad_v_dgrad_v_sum_tangent(nbk_ctx *ctx, const double *d_packed, const double *d_aux,
                                       int n_total, int K, double eps2, int t0, int t1, int s0,
                                       int s1, double *d_gsum_tan, double *d_dgsum_tan,
                                       void *stream) {
  DEV_COMMON();
  if (K <= 0 || !d_aux || ((uintptr_t)d_aux & 15) || !d_gsum_tan || !d_dgsum_tan)
    return fail(ctx, NBK_ERR_ARG, "nbk_dev_grad_v_dgrad_v_sum_tangent: bad arguments");
  const size_t nt = (size_t)(t1 - t0);
  for (int k0 = 0; k0 < K;) {
    const int kt = (K - k0 >= 3) ? 3 : 1;
    SweepParams P = base_params(d_packed, eps2, t0, t1, s0, s1);
    P.aux = d_aux + (size_t)k0 * n_total * PK;
    P.aux_stride = n_total * PK;
    P.K = kt;
    P.out[0] = d_gsum_tan + (size_t)k0 * 3 * nt;
    P.out[1] = d_dgsum_tan + (size_t)k0 * 3 * nt;
    if (kt == 3) TRY((run_tangent<3, true>(ctx, P, (cudaStream_t)stream)));
    else TRY((run_tangent<1, true>(ctx, P, (cudaStream_t)stream)));
    k0 += kt;
  }
  return NBK_OK;
}

// ---- peer-memory exchange ---------------------------------------------------------------------
int nbk_peer_alloc(nbk_ctx *ctx, size_t bytes, void **d_ptr, unsigned char *handle) {
  if (!ctx || !d_ptr || bytes == 0) return ctx ? fail(ctx, NBK_ERR_ARG, "nbk_peer_alloc: bad arguments") : NBK_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess)
    return fail(ctx, NBK_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  CK(cudaMemset(p, 0, bytes));
  if (handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) == NBK_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
      cudaFree(p);
      return fail(ctx, NBK_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof h);
  }
  *d_ptr = p;
  return NBK_OK;
}

int nbk_peer_open(nbk_ctx *ctx, const unsigned char *handle, void **d_ptr) {
  if (!ctx || !handle || !d_ptr) return ctx ? fail(ctx, NBK_ERR_ARG, "nbk_peer_open: bad arguments") : NBK_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return NBK_OK;
}

int nbk_peer_close(nbk_ctx *ctx, void *d_ptr) {
  if (!ctx) return NBK_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (d_ptr) CK(cudaIpcCloseMemHandle(d_ptr));
  return NBK_OK;
}

int nbk_peer_free(nbk_ctx *ctx, void *d_ptr) {
  if (!ctx) return NBK_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (d_ptr) CK(cudaFree(d_ptr));
  return NBK_OK;
}

int nbk_dev_publish(nbk_ctx *ctx, int nranks, int self, uint64_t *const *d_ready_of_rank,
                    uint64_t epoch, void *stream) {
  if (!ctx) return NBK_ERR_ARG;
  if (nranks < 1 || nranks > MAX_PEERS || self < 0 || self >= nranks || !d_ready_of_rank)
    return fail(ctx, NBK_ERR_ARG, "nbk_dev_publish: bad arguments");
  CK(cudaSetDevice(ctx->device));
  PublishParams F;
  memset(&F, 0, sizeof F);
  for (int r = 0; r < nranks; ++r) {
    if (!d_ready_of_rank[r]) return fail(ctx, NBK_ERR_ARG, "nbk_dev_publish: null ready array");
    F.ready_of[r] = (unsigned long long *)d_ready_of_rank[r];
  }
  publish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(F, nranks, self, (unsigned long long)epoch);
  CK(cudaGetLastError());
  ctx->launches++;
  return NBK_OK;
}

namespace {
// Validates a shard table and turns it into sweep parameters (targets must be rows of `self`).
int peer_params(nbk_ctx *ctx, const nbk_shards *sh, int n_total, double eps2, int t0, int t1,
                int s0, int s1, bool need_aux, SweepParams &P) {
  TRY(check_ranges(ctx, n_total, t0, t1, s0, s1));
  if (!sh || sh->nranks < 1 || sh->nranks > MAX_PEERS || sh->self < 0 || sh->self >= sh->nranks ||
      sh->shard_rows <= 0 || sh->shard_rows % TJ != 0 || s0 % TJ != 0)
    return fail(ctx, NBK_ERR_ARG,
                "bad shard table (1..%d ranks, shard_rows and s0 multiples of %d)", MAX_PEERS, TJ);
  if ((long long)sh->shard_rows * sh->nranks < n_total)
    return fail(ctx, NBK_ERR_ARG, "shards hold %lld rows < n_total = %d",
                (long long)sh->shard_rows * sh->nranks, n_total);
  const long long lo = (long long)sh->self * sh->shard_rows;
  if (t1 > t0 && (t0 < lo || t1 > lo + sh->shard_rows))
    return fail(ctx, NBK_ERR_ARG, "targets [%d,%d) are not rows of rank %d's shard", t0, t1,
                sh->self);
  P = base_params(nullptr, eps2, t0, t1, s0, s1);
  for (int r = 0; r < sh->nranks; ++r) {
    if (!sh->packed[r] || ((uintptr_t)sh->packed[r] & 15) ||
        (need_aux && (!sh->aux[r] || ((uintptr_t)sh->aux[r] & 15))))
      return fail(ctx, NBK_ERR_ARG, "shard %d: null or misaligned device pointer", r);
    P.shard[r] = sh->packed[r];
    P.aux_shard[r] = need_aux ? sh->aux[r] : nullptr;
  }
  P.shard_rows = sh->shard_rows;
  P.self_rank = sh->self;
  P.ready = (const unsigned long long *)sh->ready;
  P.epoch = sh->epoch;
  P.wait_ns = ctx->peer_wait_ns;
  // this rank's own shard, biased so that packed + i*PK is target i (never dereferenced outside it)
  P.packed = sh->packed[sh->self] - (size_t)lo * PK;
  if (need_aux) P.aux = sh->aux[sh->self] - (size_t)lo * PK;
  CK(cudaSetDevice(ctx->device));
  return NBK_OK;
}
}  // namespace

int nbk_dev_grad_v_dgrad_v_sum_peer(nbk_ctx *ctx, const nbk_shards *sh, int n_total, double eps2,
                                    int t0, int t1, int s0, int s1, double *d_gsum,
                                    double *d_dgsum, int accumulate, void *stream) {
  SweepParams P;
  TRY(peer_params(ctx, sh, n_total, eps2, t0, t1, s0, s1, false, P));
  if (t1 == t0 || s1 == s0) return NBK_OK;
  if (!d_gsum || !d_dgsum) return fail(ctx, NBK_ERR_ARG, "null output");
  P.out[0] = d_gsum;
  P.out[1] = d_dgsum;
  P.accumulate = accumulate;
  return run_op_peer<OpGradVDGradVB<1>>(ctx, P, (cudaStream_t)stream, false);
}

int nbk_dev_grad_v_sum_peer(nbk_ctx *ctx, const nbk_shards *sh, int n_total, double eps2, int t0,
                            int t1, int s0, int s1, double *d_gsum, int accumulate,
                            void *stream) {
  SweepParams P;
  TRY(peer_params(ctx, sh, n_total, eps2, t0, t1, s0, s1, false, P));
  if (t1 == t0 || s1 == s0) return NBK_OK;
  if (!d_gsum) return fail(ctx, NBK_ERR_ARG, "null output");
  P.out[0] = d_gsum;
  P.accumulate = accumulate;
  return run_op_peer<OpGradV>(ctx, P, (cudaStream_t)stream, false);
}

int nbk_dev_v_sum_peer(nbk_ctx *ctx, const nbk_shards *sh, int n_total, double eps2, int t0,
                       int t1, int s0, int s1, double *d_phi, int accumulate, void *stream) {
  SweepParams P;
  TRY(peer_params(ctx, sh, n_total, eps2, t0, t1, s0, s1, false, P));
  if (t1 == t0 || s1 == s0) return NBK_OK;
  if (!d_phi) return fail(ctx, NBK_ERR_ARG, "null output");
  P.out[0] = d_phi;
  P.accumulate = accumulate;
  return run_op_peer<OpV>(ctx, P, (cudaStream_t)stream, false);
}

int nbk_dev_kick_sum_peer(nbk_ctx *ctx, const nbk_shards *sh, int n_total, double eta, double dt,
                          int t0, int t1, int s0, int s1, double *d_dv, double *d_tmin,
                          void *stream) {
  SweepParams P;
  TRY(peer_params(ctx, sh, n_total, 0.0, t0, t1, s0, s1, false, P));
  if (t1 == t0 || s1 == s0) return NBK_OK;
  if (!d_dv) return fail(ctx, NBK_ERR_ARG, "null output");
  P.dt = dt;
  P.eta = eta;
  P.out[0] = d_dv;
  P.out[1] = d_tmin;
  return d_tmin ? run_op_peer<OpKick<true>>(ctx, P, (cudaStream_t)stream, false)
                : run_op_peer<OpKick<false>>(ctx, P, (cudaStream_t)stream, false);
}

int nbk_dev_grad_v_dgrad_v_sum_tangent_peer(nbk_ctx *ctx, const nbk_shards *sh, int n_total, int K,
                                            double eps2, int t0, int t1, int s0, int s1,
                                            double *d_gsum_tan, double *d_dgsum_tan,
                                            void *stream) {
  SweepParams P0;
  TRY(peer_params(ctx, sh, n_total, eps2, t0, t1, s0, s1, true, P0));
  if (t1 == t0 || s1 == s0) return NBK_OK;
  if (K <= 0 || !d_gsum_tan || !d_dgsum_tan)
    return fail(ctx, NBK_ERR_ARG, "nbk_dev_grad_v_dgrad_v_sum_tangent_peer: bad arguments");
  const size_t nt = (size_t)(t1 - t0);
  const size_t dir = (size_t)sh->shard_rows * PK;  // doubles per direction inside a shard
  for (int k0 = 0; k0 < K;) {
    const int kt = (K - k0 >= 3) ? 3 : 1;
    SweepParams P = P0;
    P.aux = P0.aux + (size_t)k0 * dir;
    for (int r = 0; r < sh->nranks; ++r) P.aux_shard[r] = P0.aux_shard[r] + (size_t)k0 * dir;
    P.aux_stride = (int)dir;
    P.K = kt;
    P.out[0] = d_gsum_tan + (size_t)k0 * 3 * nt;
    P.out[1] = d_dgsum_tan + (size_t)k0 * 3 * nt;
    if (kt == 3) TRY((run_tangent<3, true, true>(ctx, P, (cudaStream_t)stream)));
    else TRY((run_tangent<1, true, true>(ctx, P, (cudaStream_t)stream)));
    k0 += kt;
  }
  return NBK_OK;
}

int nbk_set_peer_timeout(nbk_ctx *ctx, double seconds) {
  if (!ctx || !(seconds >= 0.0) || seconds > 1e9)
    return ctx ? fail(ctx, NBK_ERR_ARG, "nbk_set_peer_timeout: seconds must be in [0, 1e9]") : NBK_ERR_ARG;
  ctx->peer_wait_ns = (unsigned long long)(seconds * 1e9);
  for (nbk_ctx *peer : ctx->peers) peer->peer_wait_ns = ctx->peer_wait_ns;
  return NBK_OK;
}

int nbk_tune(nbk_ctx *ctx, int variant) {
  if (!ctx) return NBK_ERR_ARG;
  if (variant < 0 || variant > 3) return fail(ctx, NBK_ERR_ARG, "nbk_tune: no shape %d", variant);
  ctx->shape = variant;
  return NBK_OK;
}

int nbk_fp64_peak(nbk_ctx *ctx, int iters, double *tflops, float *ms) {
  if (!ctx || !tflops || !ms || iters <= 0) return NBK_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  TRY(reserve(ctx, ctx->d_scalar, 64));
  const int blocks = ctx->sm_count * 8, threads = 256;
  dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(16, 1.0, (double *)ctx->d_scalar.p);
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(iters, 1.0, (double *)ctx->d_scalar.p);
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaGetLastError());
  CK(cudaEventSynchronize(ctx->ev1));
  CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  ctx->ev_valid = false;
  ctx->launches += 2;
  double flop = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
  *tflops = flop / ((double)*ms * 1e-3) / 1e12;
  return NBK_OK;
}

}  // extern "C"
