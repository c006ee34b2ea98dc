(* nbk.ml — OCaml binding of libnbk (include/nbk.h) for farr/direct-nbody.

   NOT COMPILED HERE (no OCaml toolchain in the build image); see nbk_stubs.c.  Drop this file
   and nbk_stubs.c into the reference tree, add Nbk to nbody.mllib, and link with
   -lnbk (INTEGRATION.md has the myocamlbuild.ml lines).

   [range] is (t0, t1, s0, s1): targets [t0,t1) x sources [s0,s1), indices into the body
   array; outputs are indexed from t0.  Every function raises [Failure msg] on error. *)

open Bigarray

type ctx
type vec = (float, float64_elt, c_layout) Array1.t
type range = int * int * int * int

external create : int -> ctx = "nbk_ml_create"

(* [create_multi ndev]: one context driving GPUs 0..ndev-1 from this single process; the
   host-buffer sweeps shard their targets over them (nbk_create_multi). *)
external create_multi : int -> ctx = "nbk_ml_create_multi"

external grad_v_sum : ctx -> vec -> vec -> float -> range -> vec -> unit
  = "nbk_ml_grad_v_sum_bc" "nbk_ml_grad_v_sum"

external grad_v_dgrad_v_sum : ctx -> vec -> vec -> vec -> float -> range -> vec -> vec -> unit
  = "nbk_ml_grad_v_dgrad_v_sum_bc" "nbk_ml_grad_v_dgrad_v_sum"

external grad_v_dgrad_v_sum_predicted :
  ctx -> vec -> vec -> vec -> vec -> vec -> vec -> float -> float -> range -> vec -> vec -> unit
  = "nbk_ml_predicted_bc" "nbk_ml_predicted"

external energy : ctx -> vec -> vec -> vec -> float -> float * float = "nbk_ml_energy"

external v_sum : ctx -> vec -> vec -> float -> range -> vec -> float
  = "nbk_ml_v_sum_bc" "nbk_ml_v_sum"

external kick_sum : ctx -> vec -> vec -> vec -> float -> float -> range -> vec -> vec -> unit
  = "nbk_ml_kick_sum_bc" "nbk_ml_kick_sum"

external grad_v_dgrad_v_sum_tangent :
  ctx -> vec -> vec -> vec -> int -> vec -> vec -> float -> range -> vec -> vec -> unit
  = "nbk_ml_tangent_bc" "nbk_ml_tangent"

(* Bodies resident on the device for Omelyan_advancer.OA (the nbk_oa_ entry points): upload once, run whole
   steps of B A C A B on the GPU, read diagnostics / the state back when needed. *)
external oa_upload : ctx -> vec -> vec -> vec -> unit = "nbk_ml_oa_upload"
external oa_steps : ctx -> float -> float -> int -> unit = "nbk_ml_oa_steps"
external oa_download : ctx -> vec -> vec -> unit = "nbk_ml_oa_download"
(* [oa_diagnostics ctx eps2 el] = (ke, pe); fills el (4n: E_i, E_i/m, |L_i|, |L_i|/m) *)
external oa_diagnostics : ctx -> float -> vec -> float * float = "nbk_ml_oa_diagnostics"

(* Advancer.A (advancer.ml:207-428).  [interact_sum ctx m q na eps2 s]: bodies [imin, n)
   re-indexed from 0, the first na active; s (3 per body) receives the grad_v sums of
   interact_bodies' pair loop (advancer.ml:213-235). *)
external interact_sum : ctx -> vec -> vec -> int -> float -> vec -> unit
  = "nbk_ml_interact_sum_bc" "nbk_ml_interact_sum"

(* The body records resident on the device, 35 doubles per body in the field order of
   include/nbk.h: t m q(3) p(3) h hmax t0 dVdq0(3) dVdq1(3) dVdq2(3) p0(3) q0(3) q1(3) q2(3) cs(3) *)
type perm = (int32, int32_elt, c_layout) Array1.t
external adv_upload : ctx -> vec -> unit = "nbk_ml_adv_upload"
external adv_download : ctx -> int -> int -> vec -> unit = "nbk_ml_adv_download"
external adv_permute : ctx -> perm -> unit = "nbk_ml_adv_permute"
external adv_begin_step : ctx -> int -> int -> float -> unit = "nbk_ml_adv_begin_step"
external adv_interact : ctx -> int -> int -> float -> float -> float -> unit
  = "nbk_ml_adv_interact_bc" "nbk_ml_adv_interact"
external adv_finish : ctx -> int -> int -> float -> vec -> unit = "nbk_ml_adv_finish"
external adv_get_t : ctx -> int -> float = "nbk_ml_adv_get_t"
external adv_init_first : ctx -> float -> unit = "nbk_ml_adv_init_first"
external adv_range_max : unit -> int = "nbk_ml_adv_range_max"
(* [adv_advance_range ctx dt sf eps2 hmax perm]: advance_range bs imax dt sf None
   (advancer.ml:316-350) ENTIRELY on the device for the prefix of imax = dim hmax rows; hmax: in,
   the prefix's hmax, out, the new hmax in the new order; perm: out, new row s = old row perm.{s} *)
external adv_advance_range : ctx -> float -> float -> float -> vec -> perm -> unit
  = "nbk_ml_adv_advance_range_bc" "nbk_ml_adv_advance_range"

(* Leapfrog with the bodies resident on the device (leapfrog.ml:106-158) *)
external lf_upload : ctx -> vec -> vec -> vec -> vec -> unit = "nbk_ml_lf_upload"
external lf_download : ctx -> vec -> vec -> vec -> unit = "nbk_ml_lf_download"
(* [lf_advance_subsystem ctx dt eta t dt_max perm]: advance_subsystem dt eta bs iend
   (leapfrog.ml:119-142) ENTIRELY on the device for iend = dim t <= nbk_lf_subsystem_max bodies *)
external lf_advance_subsystem : ctx -> float -> float -> vec -> vec -> perm -> unit
  = "nbk_ml_lf_advance_subsystem_bc" "nbk_ml_lf_advance_subsystem"

(* ---- round 2: the remaining integrator-facing entry points ------------------------------------ *)
type ivec = (int32, int32_elt, c_layout) Array1.t

external device_count : ctx -> int = "nbk_ml_device_count"

(* Energy.total_kinetic_energy alone: no potential sweep runs (pe = NULL in nbk_energy) *)
external kinetic_energy : ctx -> vec -> vec -> float = "nbk_ml_kinetic_energy"
external potential_energy : ctx -> vec -> vec -> float -> float = "nbk_ml_potential_energy"

(* Leapfrog.kick loops.  [kick_dv_sum ctx m q dt range dv]: velocities only (eta = infinity);
   [kick_block_sum ctx m q v ibegin eta dt dv tmin]: the whole loop of advance_subsystem
   (leapfrog.ml:132-136) for bodies [0, dim m) with the active block [ibegin, dim m) *)
external kick_dv_sum : ctx -> vec -> vec -> float -> range -> vec -> unit
  = "nbk_ml_kick_dv_sum_bc" "nbk_ml_kick_dv_sum"
external kick_block_sum : ctx -> vec -> vec -> vec -> int -> float -> float -> vec -> vec -> unit
  = "nbk_ml_kick_block_sum_bc" "nbk_ml_kick_block_sum"

(* DFloatBase.grad_v tangents (dFloat_advancer.ml:57-64), k directions, dq direction major *)
external grad_v_sum_tangent : ctx -> vec -> vec -> int -> vec -> float -> range -> vec -> unit
  = "nbk_ml_grad_v_sum_tangent_bc" "nbk_ml_grad_v_sum_tangent"

(* DFloat_hermite.advance_body's loop (dFloat_hermite.ml:115-135):
   [grad_v_dgrad_v_sum_predicted_tangent ctx (t, m, q, p, f, df) k (t_tan, q_tan, p_tan, f_tan, df_tan)
      nt nt_tan eps2 range (gsum, dgsum, gsum_tan, dgsum_tan)] *)
external grad_v_dgrad_v_sum_predicted_tangent :
  ctx -> vec * vec * vec * vec * vec * vec -> int -> vec * vec * vec * vec * vec -> float -> vec ->
  float -> range -> vec * vec * vec * vec -> unit
  = "nbk_ml_predicted_tangent_bc" "nbk_ml_predicted_tangent"

(* Analysis.binaries / binaries_threshold / tightest_binary (analysis.ml:836-876) *)
external binary_scan :
  ctx -> vec -> vec -> vec -> float -> float -> range -> vec -> ivec -> ivec -> unit
  = "nbk_ml_binary_scan_bc" "nbk_ml_binary_scan"
external binary_list : ctx -> vec -> vec -> vec -> float -> float -> ivec -> ivec -> vec -> int
  = "nbk_ml_binary_list_bc" "nbk_ml_binary_list"

(* packed bodies resident on the device: small-active-set callers *)
external bodies_upload : ctx -> vec -> vec -> vec -> bool -> unit = "nbk_ml_bodies_upload"
external bodies_update : ctx -> int -> int -> vec -> vec -> bool -> unit
  = "nbk_ml_bodies_update_bc" "nbk_ml_bodies_update"
external resident_count : ctx -> int = "nbk_ml_resident_count"
external resident_grad_v_sum : ctx -> float -> range -> vec -> unit = "nbk_ml_resident_grad_v_sum"
external resident_grad_v_dgrad_v_sum : ctx -> float -> range -> vec -> vec -> unit
  = "nbk_ml_resident_grad_v_dgrad_v_sum"
external resident_v_sum : ctx -> float -> range -> vec -> float = "nbk_ml_resident_v_sum"
external resident_kick_sum : ctx -> float -> float -> range -> vec -> vec -> unit
  = "nbk_ml_resident_kick_sum_bc" "nbk_ml_resident_kick_sum"
(* [resident_leapfrog_steps ctx dt nsteps q v]: nsteps DKD steps on the device (leapfrog.ml:160-165) *)
external resident_leapfrog_steps : ctx -> float -> int -> vec -> vec -> unit
  = "nbk_ml_resident_leapfrog_steps"

(* Leapfrog per-phase calls on the resident array: the recursion above lf_subsystem_max bodies *)
external lf_permute : ctx -> ivec -> unit = "nbk_ml_lf_permute"
(* [lf_sort_sub ctx perm]: sort_sub bs (dim perm) on the device; perm: out, new row s = old row perm.{s} *)
external lf_sort_sub : ctx -> ivec -> unit = "nbk_ml_lf_sort_sub"
external lf_begin : ctx -> int -> int -> unit = "nbk_ml_lf_begin"
external lf_drift : ctx -> int -> int -> float -> unit = "nbk_ml_lf_drift"
(* [lf_kick_block ctx iend ibegin eta dt with_tscale drift_after dt_max] *)
external lf_kick_block : ctx -> int -> int -> float -> float -> bool -> float -> vec -> unit
  = "nbk_ml_lf_kick_block_bc" "nbk_ml_lf_kick_block"
external lf_subsystem_max : unit -> int = "nbk_ml_lf_subsystem_max"

(* Hermite.b array resident on the device (hermite.ml:95-130, 171-178) *)
external hermite_upload : ctx -> vec -> vec -> vec -> vec -> vec -> vec -> vec -> unit
  = "nbk_ml_hermite_upload_bc" "nbk_ml_hermite_upload"
external hermite_resident_max : unit -> int = "nbk_ml_hermite_resident_max"
(* [hermite_advance_resident ctx stop_time eta eps2 max_steps] = advance_body calls made *)
external hermite_advance_resident : ctx -> float -> float -> float -> int -> int
  = "nbk_ml_hermite_advance_resident"
external hermite_download : ctx -> vec -> vec -> vec -> vec -> vec -> vec -> unit
  = "nbk_ml_hermite_download_bc" "nbk_ml_hermite_download"
(* [hermite_body_sum ctx i nt eps2 upd upd_state gsum dgsum] *)
external hermite_body_sum : ctx -> int -> float -> float -> int -> vec -> vec -> vec -> unit
  = "nbk_ml_hermite_body_sum_bc" "nbk_ml_hermite_body_sum"

external adv_energy : ctx -> float -> float * float = "nbk_ml_adv_energy"
external set_peer_timeout : ctx -> float -> unit = "nbk_ml_set_peer_timeout"
(* Pair-once sweeps (every unordered pair evaluated once, as Hermite.start_bs does,
   hermite.ml:157-166) for full squares: 0 never, 1 automatic (default), 2 whenever possible. *)
external set_sym : ctx -> int -> unit = "nbk_ml_set_sym"

let ivec n = Array1.create int32 c_layout n

(* One context per process, created on first use (the reference is single-threaded and keeps
   its configuration in globals such as Base.eps2, base.ml:20-34). *)
let the_ctx = lazy (create 0)
let ctx () = Lazy.force the_ctx

let vec n = Array1.create float64 c_layout n

(* Packing helpers: the reference keeps bodies as records of 3-element float arrays
   (SURVEY.md §8 a16); O(N) copies either side of an O(N^2) sweep. *)
let pack3 get bs =
  let n = Array.length bs in
  let a = vec (3 * n) in
  Array.iteri (fun i b -> let v = get b in
                for k = 0 to 2 do a.{3 * i + k} <- v.(k) done) bs;
  a

let pack1 get bs =
  let n = Array.length bs in
  let a = vec n in
  Array.iteri (fun i b -> a.{i} <- get b) bs;
  a
