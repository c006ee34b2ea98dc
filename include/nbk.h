/* nbk.h — C ABI of libnbk, the B200 (sm_100a) implementation of farr/direct-nbody's O(N^2)
 * pairwise-gravity hot path.
 *
 * The reference (100 % OCaml, /root/reference) has no FFI boundary: its pair functions
 * Base.v / Base.grad_v / Base.grad_v_dgrad_v (base.ml:60-110) and Leapfrog.kick
 * (leapfrog.ml:70-104) are called once per pair from double loops inside each integrator.
 * A per-pair FFI is useless for a GPU, so this ABI hoists the boundary to the LOOP: every
 * entry point below replaces one of the reference's pair loops and is what an OCaml
 * `external` stub over Bigarray.Array1 (float64, c_layout) binds (INTEGRATION.md shows the
 * stubs).  Plain pointers and sizes only; no CUDA, torch or C++ types.
 *
 * Array conventions (all caller-owned, contiguous, IEEE binary64):
 *   m[n]   masses            (B.m b)
 *   q[3n]  positions,  q[3*i+k] = (B.q b_i).(k)
 *   p[3n]  momenta,    p[3*i+k] = (B.p b_i).(k)      (Leapfrog: v[3n] velocities)
 * A sweep works on TARGETS i in [t0,t1) and SOURCES j in [s0,s1), both index ranges into the
 * same body array; a body never interacts with itself (excluded BY INDEX, so eps2 = 0 works
 * and two distinct coincident bodies still give inf/NaN like the reference).  Outputs are
 * indexed from the first target: out[3*(i-t0)+k].
 *
 * Sums are accumulated on the GPU in a fixed, run-to-run reproducible order that differs from
 * the reference's ascending-j order; results agree with the reference arithmetic to <= 1e-12
 * relative per body (tests/), not bit for bit.  Exception: nbk_kick_sum's tmin, a min over
 * pair values computed with the reference's exact IEEE operation sequence, is bit-exact
 * against the pre-kick-velocity evaluation of leapfrog.ml:94-104.
 *
 * Errors: every function returns NBK_OK (0) or a negative code; nbk_last_error gives text.
 * An OCaml stub raises Failure (nbk_last_error ctx) on non-zero (the reference's error
 * convention is exceptions).  The library never retains caller pointers past
 * a call, never calls back, and assumes one calling thread per context.
 * There is NO CPU fallback: without a CUDA device nbk_create fails with NBK_ERR_CUDA.
 */
#ifndef NBK_H
#define NBK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBK_OK 0
#define NBK_ERR_ARG (-1)   /* bad argument (null pointer, bad range, n too large) */
#define NBK_ERR_CUDA (-2)  /* CUDA runtime error (text in nbk_last_error) */
#define NBK_ERR_NOMEM (-3) /* host or device allocation failed */
#define NBK_ERR_STATE (-4) /* call sequence error (e.g. no resident bodies uploaded) */

typedef struct nbk_ctx nbk_ctx;

/* ---- context ---------------------------------------------------------------------------- */

/* Creates a context bound to CUDA device `device` (ordinal).  Owns one stream, all device
 * workspaces and pinned staging buffers; they grow on demand. */
int nbk_create(nbk_ctx **ctx, int device);
/* Single-process multi-GPU context over ndev devices (devices[] = their ordinals, or NULL for
 * 0..ndev-1) - what a single OCaml process uses to drive the GPUs of one box.  It behaves like
 * a context on the first device, except that the host-buffer sweeps nbk_grad_v_dgrad_v_sum,
 * nbk_grad_v_sum, nbk_v_sum, nbk_energy, nbk_kick_sum and the tangent sums
 * nbk_grad_v_dgrad_v_sum_tangent / nbk_grad_v_sum_tangent (BASELINE configs[4] from one process)
 * shard their TARGETS over all the devices: every
 * device uploads and packs ITS OWN rows (the host->device copies go out over all PCIe links
 * side by side), sweeps its slice over all sources - reading the other devices' rows in place
 * over NVLink (peer access, the PEER sweep kernels; no broadcast) - and the slices are read
 * back into place.  Devices that cannot map each other fall back to staging on the first device
 * and a cudaMemcpyPeerAsync broadcast.  The full acc+jerk square (targets = sources = all
 * bodies) of a large system evaluates every unordered pair once across the devices instead
 * (see the pair-once sweep below): block pairs dealt out per device, the devices' sums of a
 * slice added in device order over peer reads.  Results depend on ndev only at rounding level
 * (where a target's source range is cut into runs, or which device summed which block pairs). */
int nbk_create_multi(nbk_ctx **ctx, int ndev, const int *devices);
int nbk_device_count(const nbk_ctx *ctx);
void nbk_destroy(nbk_ctx *ctx);
/* Text of the last error on ctx; with ctx == NULL, of the last failed nbk_create. */
const char *nbk_last_error(const nbk_ctx *ctx);
const char *nbk_version(void);
/* Number of SMs, and kernel launches issued by this context so far (bench.py's gpu_launches). */
int nbk_sm_count(const nbk_ctx *ctx);
int64_t nbk_launch_count(const nbk_ctx *ctx);
/* Device time in ms of the most recent sweep's dominant kernel (CUDA events on the context's
 * stream, recorded around the sweep kernel + fix-up only). */
int nbk_last_kernel_ms(nbk_ctx *ctx, float *ms);

/* ---- host-buffer sweeps: one call = one reference pair loop ------------------------------- */

/* gsum[3*(i-t0)+k] = sum_{j in [s0,s1), j != i} (Base.grad_v m_i q_i m_j q_j).(k)
 * Replaces the Base.grad_v loops of Omelyan_advancer.OA.b_operate / c_operate
 * (omelyan_advancer.ml:44-56, 67-94) and Advancer.A.interact_bodies (advancer.ml:207-236):
 * grad_v is antisymmetric bit for bit, so the reference's symmetric i<j loop equals these
 * per-target sums. */
int nbk_grad_v_sum(nbk_ctx *ctx, int n, const double *m, const double *q, double eps2, int t0,
                   int t1, int s0, int s1, double *gsum);

/* gsum, dgsum = sum_{j != i} Base.grad_v_dgrad_v m_i q_i p_i m_j q_j p_j  (base.ml:91-110).
 * Replaces Hermite.start_bs' pair loop (hermite.ml:154-167: f_i = -gsum_i, df_i = -dgsum_i)
 * and Advancer.A.initialize_before_first_step's (advancer.ml:369-393).  THE METRIC KERNEL. */
int nbk_grad_v_dgrad_v_sum(nbk_ctx *ctx, int n, const double *m, const double *q,
                           const double *p, double eps2, int t0, int t1, int s0, int s1,
                           double *gsum, double *dgsum);

/* phi[i-t0] = sum_{j in [s0,s1), j != i} Base.v m_i q_i m_j q_j  (base.ml:60-63); the `pe`
 * accumulator of Analysis.bodies_to_el_phase_space (analysis.ml:909-914).  phi may be NULL.
 * If total != NULL, *total = sum_i phi[i] (so Energy.total_potential_energy, energy.ml:66-73,
 * is 0.5 * total for t = s = [0,n)). */
int nbk_v_sum(nbk_ctx *ctx, int n, const double *m, const double *q, double eps2, int t0, int t1,
              int s0, int s1, double *phi, double *total);

/* Energy.Make(B).total_kinetic_energy / total_potential_energy (energy.ml:57-73): *ke and
 * *pe (either may be NULL; p may be NULL when ke is).  energy = *ke + *pe (energy.ml:75-78). */
int nbk_energy(nbk_ctx *ctx, int n, const double *m, const double *q, const double *p,
               double eps2, double *ke, double *pe);

/* Leapfrog.kick (leapfrog.ml:70-104) summed over sources, every pair evaluated from the
 * velocities at call time ("pre-kick" order).  For target i:
 *   dv[3*(i-t0)+k] = sum_j (m_j * (dt / r^3)) * (q_j - q_i).(k)     (unsoftened, :83-90)
 *   tmin[i-t0]     = min_j tscale(i,j)  (leapfrog.ml:94-104; NaN ignored; +inf if no source)
 * tmin may be NULL (and eta is then unused): the eta = infinity fast path of
 * Leapfrog.advance_const_dt (leapfrog.ml:160-165).  Replaces the kick loops
 * leapfrog.ml:132-136 and :151-155. */
int nbk_kick_sum(nbk_ctx *ctx, int n, const double *m, const double *q, const double *v,
                 double eta, double dt, int t0, int t1, int s0, int s1, double *dv,
                 double *tmin);

/* Analysis.binary_energy (analysis.ml:809-822) scanned over sources, the O(N^2) loop of
 * Analysis.binaries / binaries_threshold / tightest_binary (analysis.ml:836-876).  For target i:
 *   emin[i-t0]    = min_j e(i,j)      (+inf if there is no source)
 *   partner[i-t0] = the j attaining it (lowest j on ties; -1 if none)
 *   count[i-t0]   = #{ j : e(i,j) < threshold }
 * e is evaluated with the reference's IEEE operation sequence, so comparisons and minima are
 * bit-exact; any output pointer may be NULL.  binaries uses threshold 0, binaries_threshold
 * -|eg|; tightest_binary is the global minimum of emin over targets with emin < 0. */
int nbk_binary_scan(nbk_ctx *ctx, int n, const double *m, const double *q, const double *p,
                    double eps2, double threshold, int t0, int t1, int s0, int s1, double *emin,
                    int *partner, int *count);
/* The pairs themselves: every (i, j > i) with e(i,j) < threshold, in the reference's loop
 * order (i ascending, then j), at most max_pairs of them; *npairs = the number found (an error
 * is returned if it exceeds max_pairs, with the first max_pairs in arrival order). */
int nbk_binary_list(nbk_ctx *ctx, int n, const double *m, const double *q, const double *p,
                    double eps2, double threshold, int max_pairs, int *pi, int *pj, double *e,
                    int *npairs);

/* The whole pair loop of Advancer.A.interact_bodies (advancer.ml:213-235) in one call, for the
 * bodies the caller ships ([imin, n) of the reference, re-indexed from 0) of which the first
 * `na` are the active block:
 *   i <  na:  S_i = sum_{j in [0,n), j != i} Base.grad_v(i, j)     (active x everybody)
 *   i >= na:  S_i = sum_{j in [0,na)}        Base.grad_v(i, j)     (passive x active)
 * so that dVdqK_i += (vC cs_i.(K)) S_i.  One upload, two sweeps, one read-back. */
int nbk_interact_sum(nbk_ctx *ctx, int n, int na, const double *m, const double *q, double eps2,
                     double *S);

/* The whole kick loop of Leapfrog.advance_subsystem (leapfrog.ml:132-136: for i in
 * [ibegin,iend), j in [0,i)) in one call, pre-kick order: for the bodies [0,iend),
 *   i >= ibegin: sources [0,iend) \ {i};   i < ibegin: sources [ibegin,iend).
 * dv[3 iend], tmin[iend] (tmin may be NULL) as in nbk_kick_sum. */
int nbk_kick_block_sum(nbk_ctx *ctx, int iend, int ibegin, const double *m, const double *q,
                       const double *v, double eta, double dt, double *dv, double *tmin);

/* Hermite.advance_body's loop (hermite.ml:106-115): every body is first predicted to time
 * `nt` (predict_position_into / predict_momentum_into, hermite.ml:53-72) from its own
 * (t,q,p,f,df), then gsum/dgsum as in nbk_grad_v_dgrad_v_sum over the predicted bodies
 * (fp = -gsum, dfp = -dgsum).  t[n], f[3n], df[3n] as in Hermite.b (hermite.ml:20-29). */
int nbk_grad_v_dgrad_v_sum_predicted(nbk_ctx *ctx, int n, const double *t, const double *m,
                                     const double *q, const double *p, const double *f,
                                     const double *df, double nt, double eps2, int t0, int t1,
                                     int s0, int s1, double *gsum, double *dgsum);

/* Tangent (first-order forward-mode) of nbk_grad_v_dgrad_v_sum for K directions:
 * DFloatBase.grad_v_dgrad_v (dFloat_advancer.ml:66-81) accumulated as in
 * DFloat_hermite.start_bs (dFloat_hermite.ml:177-186).  dq, dp are [K][3n] (direction
 * major); masses and eps2 carry no tangent (dFloat_hermite.ml:36-46).  Outputs: value sums
 * gsum, dgsum [3*nt] (either may be NULL) and tangent sums gsum_tan, dgsum_tan [K][3*nt]. */
int nbk_grad_v_dgrad_v_sum_tangent(nbk_ctx *ctx, int n, const double *m, const double *q,
                                   const double *p, int K, const double *dq, const double *dp,
                                   double eps2, int t0, int t1, int s0, int s1, double *gsum,
                                   double *dgsum, double *gsum_tan, double *dgsum_tan);

/* DFloat_hermite.advance_body's loop (dFloat_hermite.ml:115-135): the dual-number copy of
 * nbk_grad_v_dgrad_v_sum_predicted.  Every body carries K tangents of its whole Hermite state -
 * t_tan [K][n] (may be NULL: zeros), q_tan, p_tan, f_tan, df_tan [K][3n], direction major; masses,
 * eps2 and eta carry none (dFloat_hermite.ml:36-46, dFloat_pushforward.ml:51) - and the target
 * time is the dual nt + nt_tan[K] (may be NULL: zeros; next_time of the stepping body,
 * dFloat_hermite.ml:81-82).  Each body is predicted to nt over duals (predict_position_into /
 * predict_momentum_into, dFloat_hermite.ml:60-79: values by the reference's operation sequence,
 * tangents by the product and quotient rules on the same expression tree), then
 * DFloatBase.grad_v_dgrad_v (dFloat_advancer.ml:66-81) is summed over the predicted sources:
 * value sums gsum, dgsum [3 nt] (either may be NULL; fp = -gsum, dfp = -dgsum) and their
 * tangents gsum_tan, dgsum_tan [K][3 nt].  With K = 6n unit directions this is one
 * advance_body of DFloat_pushforward.jacobian_advance (dFloat_pushforward.ml:39-76). */
int nbk_grad_v_dgrad_v_sum_predicted_tangent(
    nbk_ctx *ctx, int n, const double *t, const double *m, const double *q, const double *p,
    const double *f, const double *df, int K, const double *t_tan, const double *q_tan,
    const double *p_tan, const double *f_tan, const double *df_tan, double nt,
    const double *nt_tan, double eps2, int t0, int t1, int s0, int s1, double *gsum,
    double *dgsum, double *gsum_tan, double *dgsum_tan);

/* Tangent of nbk_grad_v_sum (DFloatBase.grad_v, dFloat_advancer.ml:57-64). */
int nbk_grad_v_sum_tangent(nbk_ctx *ctx, int n, const double *m, const double *q, int K,
                           const double *dq, double eps2, int t0, int t1, int s0, int s1,
                           double *gsum, double *gsum_tan);

/* ---- device-resident bodies: small-active-set callers keep the N bodies on the GPU -------- */

/* Uploads and packs n bodies; `vel` is momenta (is_velocity = 0, the BODY convention) or
 * velocities (is_velocity = 1, the Leapfrog.b convention, leapfrog.ml:5-12). */
int nbk_bodies_upload(nbk_ctx *ctx, int n, const double *m, const double *q, const double *vel,
                      int is_velocity);
/* Overwrites bodies [i0,i1) (q and vel are [3*(i1-i0)]; masses unchanged). */
int nbk_bodies_update(nbk_ctx *ctx, int i0, int i1, const double *q, const double *vel,
                      int is_velocity);
int nbk_resident_count(const nbk_ctx *ctx);
int nbk_resident_grad_v_sum(nbk_ctx *ctx, double eps2, int t0, int t1, int s0, int s1,
                            double *gsum);
int nbk_resident_grad_v_dgrad_v_sum(nbk_ctx *ctx, double eps2, int t0, int t1, int s0, int s1,
                                    double *gsum, double *dgsum);
int nbk_resident_v_sum(nbk_ctx *ctx, double eps2, int t0, int t1, int s0, int s1, double *phi,
                       double *total);
int nbk_resident_kick_sum(nbk_ctx *ctx, double eta, double dt, int t0, int t1, int s0, int s1,
                          double *dv, double *tmin);

/* nsteps plain drift-kick-drift steps of size dt on the resident bodies (uploaded with
 * is_velocity = 1), entirely on the device: the steady state of Leapfrog.advance_const_dt
 * (leapfrog.ml:160-165 with every body active: drift dt/2, all-pairs kick, drift dt/2,
 * leapfrog.ml:129-137).  q, v (may be NULL) receive the final positions and velocities. */
int nbk_resident_leapfrog_steps(nbk_ctx *ctx, double dt, int nsteps, double *q, double *v);

/* ---- device-resident Advancer.A records (advancer.ml:143-428) -------------------------------- */
/* The Advancer.A.body array (advancer.ml:27-45) lives on the device as records[n][35] =
 * {t, m, q[3], p[3], h, hmax, t0, dVdq0[3], dVdq1[3], dVdq2[3], p0[3], q0[3], q1[3], q2[3],
 * cs[3]} in the reference's array order; the host keeps the control flow of advance_range
 * (advancer.ml:316-350: block recursion, hmax, sorts) and calls one function per step phase.
 * Every per-body formula repeats the reference's operation sequence, so results are
 * bit-identical to doing the same phases on the host around nbk_interact_sum.  Calls are
 * enqueued; only functions that return data synchronise. */
int nbk_adv_upload(nbk_ctx *ctx, int n, const double *records);
/* rows [i0, i1) -> records[(i1-i0)][35] */
int nbk_adv_download(nbk_ctx *ctx, int i0, int i1, double *records);
/* sort_range (advancer.ml:290-297): new rows [0,count) := old rows perm[0..count) */
int nbk_adv_permute(nbk_ctx *ctx, int count, const int *perm);
/* predict_q012 bs.(i) hnew; clear_before_step bs.(i)  for i in [imin, imax) (advancer.ml:329-332) */
int nbk_adv_begin_step(nbk_ctx *ctx, int imin, int imax, double hnew);
/* interact_bodies bs imin imax t vC (advancer.ml:207-236): predict_position of rows [imin,n),
 * the two rectangular grad_v sweeps, dVdqK += (vC cs.(K)) S - all on the device */
int nbk_adv_interact(nbk_ctx *ctx, int imin, int imax, double t, double vC, double eps2);
/* finish_body bs.(i) t for i in [imin, imax) (advancer.ml:238-252); if records != NULL the
 * finished rows are read back (the host needs q, q2, dVdq2, h, m for h_adaptive, whose libm
 * pow keeps the time-step sequence identical to the reference's) */
int nbk_adv_finish(nbk_ctx *ctx, int imin, int imax, double t, double *records);
int nbk_adv_get_t(nbk_ctx *ctx, int i, double *t);
/* initialize_before_first_step's all-pairs acc+jerk loop and accumulator set-up
 * (advancer.ml:357-393) on the device; the host then reads dVdq0-2 to assign the first steps */
int nbk_adv_init_first(nbk_ctx *ctx, double eps2);
/* Energy.energy of the resident bodies */
int nbk_adv_energy(nbk_ctx *ctx, double eps2, double *ke, double *pe);
/* `advance_range bs imax dt sf None` (advancer.ml:316-350) ENTIRELY on the device, for a prefix
 * of imax <= nbk_adv_range_max() rows of the resident records: one persistent cooperative kernel
 * runs the block recursion (frame stack, block search, predict_q012, the three interact_bodies
 * of every block step against ALL rows [imin, n), finish_body, h_adaptive, sort_range) with one
 * grid barrier per interact_bodies and one per sort - no host round trip per block step.  hmax[imax]: in, the
 * prefix's hmax (the host's mirror); out, the new hmax in the new row order.  perm[imax]: out,
 * new row s = old row perm[s] (the device rows have already been moved; the caller applies it to
 * whatever it keeps per row, e.g. ids).  stats (may be NULL): += {interact_bodies calls, pair
 * interactions}.  The sums are reproducible run to run but ordered differently from
 * nbk_adv_interact's, and h_adaptive uses the device's pow (<= 2 ulp) instead of libm's; hmax
 * only steers comparisons, so the block schedule is the reference's unless such a comparison
 * is decided in the last bit.  Agreement with the per-phase path: rounding (1e-13), not bits. */
int nbk_adv_advance_range(nbk_ctx *ctx, int imax, double dt, double sf, double eps2, double *hmax,
                          int *perm, long long *stats);
int nbk_adv_range_max(void);
/* Diagnostics of nbk_adv_advance_range since the last reset: prof[0] block steps, prof[1..9]
 * clock cycles CTA 0 spent in {control, deferred folds + active-block prediction, predictors,
 * slot x active sums, active x own-tile sums, barrier, fold + finish_body + h_adaptive, barrier,
 * sort}, prof[10] kernel launches. */
int nbk_adv_range_profile(nbk_ctx *ctx, long long *prof, int reset);

/* ---- device-resident Leapfrog.b array (leapfrog.ml:106-158) ---------------------------------- */
/* The bodies of Leapfrog.advance / advance_subsystem stay on the device in the reference's array
 * order; the host keeps the block recursion (leapfrog.ml:119-142) - which needs nothing but
 * dt_max - and calls one function per phase.  Every formula repeats the reference's operation
 * sequence, so results are bit-identical to the same phases on the host around
 * nbk_kick_block_sum.  Calls are enqueued; only functions that return data synchronise. */
int nbk_lf_upload(nbk_ctx *ctx, int n, const double *dt_max, const double *m, const double *q,
                  const double *v);
/* any pointer may be NULL */
int nbk_lf_download(nbk_ctx *ctx, double *dt_max, double *q, double *v);
/* sort_sub (leapfrog.ml:106-111): new rows [0,count) := old rows perm[0..count) */
int nbk_lf_permute(nbk_ctx *ctx, int count, const int *perm);
/* sort_sub bs iend (leapfrog.ml:106-111) computed ON THE DEVICE from the resident dt_max: the
 * stable sort of rows [0, iend) by OCaml's `compare` on dt_max (nan lowest, nan = nan), rows and
 * dt_max moved accordingly; perm[iend]: out, new row s = old row perm[s] (for what the caller
 * keeps per row).  Replaces a host sort of up to n keys plus nbk_lf_permute. */
int nbk_lf_sort_sub(nbk_ctx *ctx, int iend, int *perm);
/* dt_max := infinity on [ibegin, iend)   (leapfrog.ml:125-127).  nbk_lf_kick_block performs
 * this reset itself for its active block - nothing reads those values in between - so a caller
 * following advance_subsystem need not call it. */
int nbk_lf_begin(nbk_ctx *ctx, int ibegin, int iend);
/* drift dt on bodies [i0, i1)   (leapfrog.ml:64-68; the host adds dt to its own copy of t) */
int nbk_lf_drift(nbk_ctx *ctx, int i0, int i1, double dt);
/* The kick loop of advance_subsystem (leapfrog.ml:132-136) as in nbk_kick_block_sum, applied in
 * place: v += dv; dt_max = min(dt_max, tmin), an active body [ibegin, iend) starting from
 * infinity (with_tscale = 0: velocities only, the eta = infinity case); then the drift that
 * follows the kick (leapfrog.ml:137) if with_drift != 0: drift drift_after on the active bodies.
 * dt_max (may be NULL) receives the updated dt_max[0..iend). */
int nbk_lf_kick_block(nbk_ctx *ctx, int iend, int ibegin, double eta, double dt, int with_tscale,
                      int with_drift, double drift_after, double *dt_max);

/* `advance_subsystem dt eta bs iend` (leapfrog.ml:119-142) ENTIRELY on the device, for a prefix of
 * iend <= nbk_lf_subsystem_max() resident bodies: the call touches no body beyond iend, so one
 * CTA - or, beyond 96 bodies, an 8-CTA thread-block cluster with a replica per CTA and the
 * targets of every kick block dealt out - keeps them in shared memory and runs the whole block
 * recursion (find_can_advance_index, the dt_max resets, drifts, every kick block with its dt_max
 * updates, sort_sub): one launch instead of one launch group and one read-back per kick block.  t[iend]: in/out, the bodies'
 * times (the device keeps no t).  dt_max[iend]: out, in the new array order.  perm[iend]: out,
 * new row s = old row perm[s] (the device rows have been moved).  *nkicks += kick blocks.
 * Pair time-scales are bit-exact as in nbk_lf_kick_block; the velocity sums are ordered
 * differently from the tiled sweeps' (agreement to rounding). */
int nbk_lf_advance_subsystem(nbk_ctx *ctx, int iend, double dt, double eta, double *t, double *dt_max,
                             int *perm, long long *nkicks);
int nbk_lf_subsystem_max(void);

/* ---- device-resident Omelyan_advancer.OA (omelyan_advancer.ml:44-104) ----------------------- */
/* The bodies (m, q, p) stay on the device; nbk_oa_steps runs nsteps of OA.advance h
 * (B A C A B: three grad_v sweeps at q, one at q*, omelyan_advancer.ml:96-104) without a host
 * round trip.  Every O(N) formula repeats the reference's operation sequence, so the result
 * is bit-identical to stepping through nbk_grad_v_sum from the host.  A step's first b_operate
 * reuses the sums of the previous step's last one (q does not move in between): 3 sweeps per
 * step in the steady state instead of 4, same bits. */
int nbk_oa_upload(nbk_ctx *ctx, int n, const double *m, const double *q, const double *p);
int nbk_oa_steps(nbk_ctx *ctx, double h, double eps2, int nsteps);
int nbk_oa_download(nbk_ctx *ctx, double *q, double *p);
/* Energy.energy (energy.ml:57-78) and, if el != NULL, Analysis.bodies_to_el_phase_space
 * (analysis.ml:904-919; el[4n] = {E_i, E_i/m, |L_i|, |L_i|/m}) of the resident bodies from ONE
 * potential sweep: *ke + *pe is the total energy. */
int nbk_oa_diagnostics(nbk_ctx *ctx, double eps2, double *ke, double *pe, double *el);

/* ---- device-resident Hermite state: one body steps at a time (hermite.ml:95-130, 171-178) --- */

/* Uploads the full Hermite.b array state (hermite.ml:20-29): t[n], m[n], q, p, f, df [3n] and
 * the time steps h[n] (may be NULL: zeros). */
int nbk_hermite_upload(nbk_ctx *ctx, int n, const double *t, const double *m, const double *q,
                       const double *p, const double *h, const double *f, const double *df);
/* One Hermite.advance_body force evaluation on the resident state: first overwrites body
 * `upd` (if upd >= 0) with upd_state[13] = {t, q[3], p[3], f[3], df[3]} - the result of the
 * previous step - then predicts every body to time nt (hermite.ml:53-72) and returns
 * gsum[3], dgsum[3] for target i over all other bodies (fp = -gsum, dfp = -dgsum,
 * hermite.ml:106-115).  One kernel launch + a 48-byte read-back per step. */
int nbk_hermite_body_sum(nbk_ctx *ctx, int i, double nt, double eps2, int upd,
                         const double *upd_state, double *gsum, double *dgsum);
/* Hermite.advance_bodies + finish_bs (hermite.ml:132-143, 171-178) run ENTIRELY on the device
 * by one persistent kernel over the resident state (n <= nbk_hermite_resident_max(): one CTA
 * up to 1536 bodies, an 8-CTA cluster up to 12288, the whole grid - cooperative launch - up to
 * 1600 bodies per SM; all three software-pipelined: the corrector of one step runs beside the
 * scheduler search and the force sum of the next, one barrier per step; NBK_HERMITE_SPEC=0 in the
 * environment at nbk_create selects the plain loops): steps
 * single bodies in the reference's list order until every next time exceeds stop_time, then
 * steps every body to stop_time.  *nsteps = advance_body calls made (Base.nsteps,
 * hermite.ml:124).  Stops early, returning NBK_ERR_STATE, after max_steps steps of the first
 * phase (the reference would loop forever when h underflows). */
int nbk_hermite_resident_max(void);
int nbk_hermite_advance_resident(nbk_ctx *ctx, double stop_time, double eta, double eps2,
                                 long max_steps, long *nsteps);
/* Reads the resident state back (any pointer may be NULL). */
int nbk_hermite_download(nbk_ctx *ctx, double *t, double *q, double *p, double *h, double *f,
                         double *df);

/* ---- raw device-pointer API (one rank of a multi-GPU job; bench.py / torch harness) ------- */

/* A packed body is 8 doubles {x, y, z, m, vx, vy, vz, 0} (64 B): the layout j-tiles are
 * bulk-copied into shared memory in.  All pointers are device pointers on ctx's device;
 * `stream` is a cudaStream_t cast to void* (NULL = the legacy default stream, as in
 * every CUDA API); work is enqueued, not waited for. */
#define NBK_PACKED_DOUBLES 8
int nbk_dev_pack(nbk_ctx *ctx, int n, const double *d_m, const double *d_q, const double *d_vel,
                 int is_velocity, double *d_packed, void *stream);
/* Targets are d_packed[t0..t1), sources d_packed[s0..s1) of one array of n_total packed
 * bodies (after an all-gather: the full system; targets = this rank's shard).
 * accumulate != 0 adds onto d_gsum / d_dgsum instead of overwriting (used to split one sweep
 * into a local-sources part overlapped with the all-gather and a remote-sources part). */
int nbk_dev_grad_v_dgrad_v_sum(nbk_ctx *ctx, const double *d_packed, int n_total, double eps2,
                               int t0, int t1, int s0, int s1, double *d_gsum, double *d_dgsum,
                               int accumulate, void *stream);
int nbk_dev_grad_v_sum(nbk_ctx *ctx, const double *d_packed, int n_total, double eps2, int t0,
                       int t1, int s0, int s1, double *d_gsum, int accumulate, void *stream);
int nbk_dev_v_sum(nbk_ctx *ctx, const double *d_packed, int n_total, double eps2, int t0, int t1,
                  int s0, int s1, double *d_phi, int accumulate, void *stream);

/* Leapfrog.kick sums on device pointers (packed with velocities): d_dv[3 (t1-t0)], and
 * d_tmin[t1-t0] unless NULL (the eta = infinity fast path), as nbk_kick_sum. */
int nbk_dev_kick_sum(nbk_ctx *ctx, const double *d_packed, int n_total, double eta, double dt,
                     int t0, int t1, int s0, int s1, double *d_dv, double *d_tmin, void *stream);

/* Tangent sweep on device pointers (one rank of a sharded DFloat_pushforward force evaluation,
 * BASELINE configs[4]).  d_aux is [K][n_total][8]: direction k's rows {dq, 0, dv, 0} built by
 * nbk_dev_pack_aux (dv = dp / m).  Outputs d_gsum_tan, d_dgsum_tan are [K][3 (t1-t0)]. */
int nbk_dev_pack_aux(nbk_ctx *ctx, int n, const double *d_m, const double *d_dq,
                     const double *d_dp, double *d_aux, void *stream);
/* The Hermite predictor fused with the pack, on device pointers (one rank predicts its own
 * slice before a sharded sweep): rows {qp, m, pp/m} of nbk_grad_v_dgrad_v_sum_predicted, and
 * the direction rows {dqp, 0, dpp/m, 0} of nbk_grad_v_dgrad_v_sum_predicted_tangent (direction k
 * at d_aux + k * aux_stride doubles; tangent arrays [K][n] / [K][3n]; d_t_tan and d_nt_tan may be
 * NULL). */
int nbk_dev_predict_pack(nbk_ctx *ctx, int n, const double *d_t, const double *d_m,
                         const double *d_q, const double *d_p, const double *d_f,
                         const double *d_df, double nt, double *d_packed, void *stream);
int nbk_dev_predict_pack_tangent(nbk_ctx *ctx, int n, int K, const double *d_t, const double *d_m,
                                 const double *d_q, const double *d_p, const double *d_f,
                                 const double *d_df, double nt, const double *d_t_tan,
                                 const double *d_q_tan, const double *d_p_tan,
                                 const double *d_f_tan, const double *d_df_tan,
                                 const double *d_nt_tan, double *d_aux, size_t aux_stride,
                                 void *stream);
int nbk_dev_grad_v_dgrad_v_sum_tangent(nbk_ctx *ctx, const double *d_packed, const double *d_aux,
                                       int n_total, int K, double eps2, int t0, int t1, int s0,
                                       int s1, double *d_gsum_tan, double *d_dgsum_tan,
                                       void *stream);

/* ---- peer-memory exchange: the all-gather fused into the sweep ------------------------------ */
/* One rank of a target-sharded sweep needs every body as a source.  Instead of
 * all-gathering the packed rows in front of the sweep, the *_peer sweeps read each source tile
 * straight out of the owning GPU's memory over NVLink/NVSwitch, behind the same TMA ring that
 * stages local tiles, so the exchange overlaps the arithmetic tile by tile and costs no step of
 * its own.  Body j lives in row j % shard_rows of rank j / shard_rows.
 *
 * Regions: nbk_peer_alloc gives zeroed device memory plus (if handle != NULL) a 64-byte CUDA IPC
 * handle another PROCESS passes to nbk_peer_open to map it (one process per GPU, the handles
 * travel over torch.distributed); inside one process plain device pointers work once peer
 * access is enabled (nbk_create_multi does that).
 * Ordering: a rank packs its rows, then nbk_dev_publish stores `epoch` into slot `self` of every
 * rank's ready[] array (release, system scope); a *_peer sweep given ready != NULL does not read
 * rank r's rows before ready[r] >= epoch (acquire).  With two row buffers used alternately no
 * further synchronisation is needed: a rank packs epoch e+1 only after its own sweep of epoch
 * e, which has seen every peer publish e, i.e. finish reading epoch e-1. */
#define NBK_MAX_PEERS 8
#define NBK_IPC_HANDLE_BYTES 64
typedef struct nbk_shards {
  int nranks, self;
  int shard_rows;                      /* rows per rank: a multiple of 128 */
  const double *packed[NBK_MAX_PEERS]; /* rank r's packed rows [shard_rows][8], mapped on this device */
  const double *aux[NBK_MAX_PEERS];    /* tangent sweeps: rank r's direction rows [K][shard_rows][8] */
  const uint64_t *ready;               /* this rank's ready[NBK_MAX_PEERS], or NULL: caller ordered the streams */
  uint64_t epoch;
} nbk_shards;
int nbk_peer_alloc(nbk_ctx *ctx, size_t bytes, void **d_ptr, unsigned char *handle);
int nbk_peer_open(nbk_ctx *ctx, const unsigned char *handle, void **d_ptr);
int nbk_peer_close(nbk_ctx *ctx, void *d_ptr);
int nbk_peer_free(nbk_ctx *ctx, void *d_ptr);
/* How long a *_peer sweep waits for a peer's ready flag before the kernel traps (the host then
 * sees a launch failure on its next call instead of a hung GPU): default 20 s, or the
 * environment variable NBK_PEER_TIMEOUT_S at nbk_create; 0 = wait forever (debugger, profiler,
 * ranks with long host-side work between steps).  All ranks must call their sweeps in lockstep:
 * a rank that steps fewer epochs than its peers leaves them waiting. */
int nbk_set_peer_timeout(nbk_ctx *ctx, double seconds);
/* d_ready_of_rank[r] = rank r's ready[] array as mapped on this device. */
int nbk_dev_publish(nbk_ctx *ctx, int nranks, int self, uint64_t *const *d_ready_of_rank,
                    uint64_t epoch, void *stream);
/* As nbk_dev_grad_v_dgrad_v_sum / _grad_v_sum / _v_sum / _tangent with the sources sharded;
 * targets [t0,t1) must be rows of rank `self`, s0 a multiple of 128. */
int nbk_dev_grad_v_dgrad_v_sum_peer(nbk_ctx *ctx, const nbk_shards *sh, int n_total, double eps2,
                                    int t0, int t1, int s0, int s1, double *d_gsum,
                                    double *d_dgsum, int accumulate, void *stream);
int nbk_dev_grad_v_sum_peer(nbk_ctx *ctx, const nbk_shards *sh, int n_total, double eps2, int t0,
                            int t1, int s0, int s1, double *d_gsum, int accumulate, void *stream);
int nbk_dev_v_sum_peer(nbk_ctx *ctx, const nbk_shards *sh, int n_total, double eps2, int t0,
                       int t1, int s0, int s1, double *d_phi, int accumulate, void *stream);
int nbk_dev_kick_sum_peer(nbk_ctx *ctx, const nbk_shards *sh, int n_total, double eta, double dt,
                          int t0, int t1, int s0, int s1, double *d_dv, double *d_tmin,
                          void *stream);
int nbk_dev_grad_v_dgrad_v_sum_tangent_peer(nbk_ctx *ctx, const nbk_shards *sh, int n_total, int K,
                                            double eps2, int t0, int t1, int s0, int s1,
                                            double *d_gsum_tan, double *d_dgsum_tan, void *stream);

/* Pair-once (symmetric) acc+jerk sweep, the shape of the reference's own start_bs loop
 * (/root/reference/hermite.ml:157-166: every unordered pair evaluated once, f_i -= gv,
 * f_j += gv).  The bodies are cut into blocks of 1024 rows; a work item is one block pair of the
 * upper triangle (diagonal included), nbk_sym_items(n) of them, numbered row-major.  A call
 * runs items [item0, item1) and writes d_sum6[b][0..5] = {gsum_b, dgsum_b} restricted to the
 * pairs of those items, for ALL n_total bodies (zeros where no item touched b): the sum over
 * disjoint item ranges covering [0, items) is nbk_grad_v_dgrad_v_sum's result (to summation
 * order).  A rank of a sharded run takes a contiguous range of items and the ranks' d_sum6 meet
 * in a reduce-scatter (nbk/sharded.py); _peer reads the rows from the owning GPUs
 * (shard_rows a multiple of 1024) and waits for every rank's ready flag first.
 * nbk_grad_v_dgrad_v_sum and its _resident / _dev forms take this path by themselves for a
 * full square (targets = sources) once there are enough block pairs to fill the SMs; so do the
 * acceleration-only and potential sums (nbk_grad_v_sum, nbk_v_sum, nbk_energy, the resident
 * nbk_oa_* steps and their _dev forms: /root/reference/omelyan_advancer.ml:44-94,
 * energy.ml:66-73) with their own pair-once kernels.  Results agree with the ordered sweeps to
 * summation order and are run-to-run identical.
 * nbk_set_sym: 0 = never, 1 = automatic (default; NBK_SYM in the environment at nbk_create),
 * 2 = for every full square of at least 2048 bodies. */
int nbk_set_sym(nbk_ctx *ctx, int mode);
int nbk_sym_items(int n_total, long long *items);
int nbk_dev_grad_v_dgrad_v_sym(nbk_ctx *ctx, const double *d_packed, int n_total, double eps2,
                               long long item0, long long item1, double *d_sum6, void *stream);
int nbk_dev_grad_v_dgrad_v_sym_peer(nbk_ctx *ctx, const nbk_shards *sh, int n_total, double eps2,
                                    long long item0, long long item1, double *d_sum6,
                                    void *stream);
/* nbk_dev_pack with the rows stored into nranks destinations: d_dst[r] = this rank's place in
 * rank r's gathered row array, mapped on this device.  The exchange step of the sharded
 * pair-once sweep ("each step all-gathers positions, velocities and masses"), done by the
 * packing kernel's own remote stores; nbk_dev_publish after it makes the rows visible, and the
 * sweep on every rank then reads its LOCAL gathered array behind the ready flags (an nbk_shards
 * whose packed[r] all point into that local array). */
int nbk_dev_pack_push(nbk_ctx *ctx, int n, const double *d_m, const double *d_q,
                      const double *d_vel, int is_velocity, int nranks, double *const *d_dst,
                      void *stream);
/* The reduce-scatter of the sharded pair-once sweep as the ranks' own peer reads: gsum, dgsum
 * rows of bodies [b0, b1) = the sum over ranks r (in rank order - deterministic) of
 * d_sum6_of_rank[r][body][0..5], rank r's nbk_dev_grad_v_dgrad_v_sym(_peer) output as mapped on
 * this device.  d_ready != NULL: this rank's flag array for the sums; rank r's sums are read
 * once d_ready[r] >= epoch (every rank publishes that array with nbk_dev_publish after its
 * sweep); waits are bounded by nbk_set_peer_timeout.  A rank may overwrite its sums in the next
 * sweep without a second buffer: its next sweep waits for every peer's NEXT row flags, which
 * the peers publish after this gather on their streams. */
int nbk_dev_sym_gather_peer(nbk_ctx *ctx, int nranks, int self, const double *const *d_sum6_of_rank,
                            const uint64_t *d_ready, uint64_t epoch, int b0, int b1,
                            double *d_gsum, double *d_dgsum, void *stream);
/* Rows [count][6] of (reduce-scattered) sums -> gsum, dgsum [count][3]. */
int nbk_dev_split6(nbk_ctx *ctx, const double *d_sum6, int count, double *d_gsum, double *d_dgsum,
                   void *stream);

/* Tile shape of the sweep kernels: 0 = automatic (by number of targets), 1 = large (1024
 * targets per CTA: 256 threads x 4), 2 = medium (256: 128 x 2), 3 = small (128: 128 x 1).
 * Results do not depend on the shape beyond summation order.  A knob for profiles/ and
 * tests; callers never need it. */
int nbk_tune(nbk_ctx *ctx, int variant);

/* FP64 DFMA-throughput microbenchmark on ctx's device (the measured FP64 peak the roofline
 * fraction is quoted against): runs `iters` dependent-chain-free DFMA per thread on a full
 * grid; returns TFLOP/s (2 flop per DFMA) in *tflops and the kernel time in *ms. */
int nbk_fp64_peak(nbk_ctx *ctx, int iters, double *tflops, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* NBK_H */
