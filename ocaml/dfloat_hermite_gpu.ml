(* dfloat_hermite_gpu.ml - the two pair loops of dFloat_hermite.ml over libnbk, for K tangent
   directions carried at once.  Diff.DFloat's [jacobian] pushes one unit direction per call
   through [advance]; the directions do not interact, so the K = 6n columns of
   DFloat_pushforward.jacobian_advance (dFloat_pushforward.ml:61-76) can ride together: the
   state below keeps every field's value and its K tangents in flat arrays (direction major),
   which is also the layout the device wants.  NOT COMPILED HERE; mirrored by
   direct-nbody_b200/host/dfloat_hermite.cpp, which tests/test_dfloat_gpu.py check against the
   oracle's template over dual numbers.

   [start_bs] (dFloat_hermite.ml:168-189)   -> Nbk.grad_v_dgrad_v_sum_tangent (+ value sums)
   [advance_body]'s List.iter (:126-135)    -> Nbk.grad_v_dgrad_v_sum_predicted_tangent:
     every body is predicted over duals ON THE DEVICE (tangents of t, q, p, f, df and of the
     target time nt), then DFloatBase.grad_v_dgrad_v is summed for the stepping body. *)

type dstate = {
  n : int; k : int;
  t : Nbk.vec; m : Nbk.vec; q : Nbk.vec; p : Nbk.vec; f : Nbk.vec; df : Nbk.vec;   (* values *)
  t_tan : Nbk.vec; q_tan : Nbk.vec; p_tan : Nbk.vec; f_tan : Nbk.vec; df_tan : Nbk.vec;
}

(* fp, dfp and their tangents for body [i] stepping to the dual time (nt, nt_tan):
   returns (fp, dfp, fp_tan, dfp_tan) with fp = -gsum etc. (dFloat_hermite.ml:131-134) *)
let force_on_predicted s i nt nt_tan =
  let g = Nbk.vec 3 and dg = Nbk.vec 3 and gt = Nbk.vec (3 * s.k) and dgt = Nbk.vec (3 * s.k) in
  Nbk.grad_v_dgrad_v_sum_predicted_tangent (Nbk.ctx ())
    (s.t, s.m, s.q, s.p, s.f, s.df) s.k (s.t_tan, s.q_tan, s.p_tan, s.f_tan, s.df_tan)
    nt nt_tan !Base.eps2 (i, i + 1, 0, s.n) (g, dg, gt, dgt);
  let neg a = Array.init (Bigarray.Array1.dim a) (fun j -> -. a.{j}) in
  (neg g, neg dg, neg gt, neg dgt)

(* start_bs: f = -gsum, df = -dgsum with their K tangents, all pairs *)
let start_forces s =
  let n3 = 3 * s.n in
  let gt = Nbk.vec (s.k * n3) and dgt = Nbk.vec (s.k * n3) in
  Nbk.grad_v_dgrad_v_sum_tangent (Nbk.ctx ()) s.m s.q s.p s.k s.q_tan s.p_tan !Base.eps2
    (0, s.n, 0, s.n) gt dgt;
  let g = Nbk.vec n3 and dg = Nbk.vec n3 in
  Nbk.grad_v_dgrad_v_sum (Nbk.ctx ()) s.m s.q s.p !Base.eps2 (0, s.n, 0, s.n) g dg;
  for j = 0 to n3 - 1 do s.f.{j} <- -. g.{j}; s.df.{j} <- -. dg.{j} done;
  for j = 0 to s.k * n3 - 1 do s.f_tan.{j} <- -. gt.{j}; s.df_tan.{j} <- -. dgt.{j} done
