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

(* Bodies resident on the device for Omelyan_advancer.OA (nbk_oa_*): upload once, run whole
   steps of B A C A B on the GPU, read diagnostics / the state back when needed. *)
external oa_upload : ctx -> vec -> vec -> vec -> unit = "nbk_ml_oa_upload"
external oa_steps : ctx -> float -> float -> int -> unit = "nbk_ml_oa_steps"
external oa_download : ctx -> vec -> vec -> unit = "nbk_ml_oa_download"
(* [oa_diagnostics ctx eps2 el] = (ke, pe); fills el (4n: E_i, E_i/m, |L_i|, |L_i|/m) *)
external oa_diagnostics : ctx -> float -> vec -> float * float = "nbk_ml_oa_diagnostics"

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
