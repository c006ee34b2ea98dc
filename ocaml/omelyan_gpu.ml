(* omelyan_gpu.ml — OA.b_operate and OA.c_operate (omelyan_advancer.ml:44-56, 67-94) over
   Nbk.grad_v_sum; a_operate and advance are unchanged.  grad_v is antisymmetric bit for bit, so
   the reference's  p_i -= f gv(i,j); p_j += f gv(i,j)  over pairs equals  p_i -= f gsum_i.
   NOT COMPILED HERE. *)

open Omelyan_advancer.OA

let gsum_of get bs =
  let arr = Array.of_list bs in
  let n = Array.length arr in
  let m = Nbk.pack1 (fun b -> b.m) arr and q = Nbk.pack3 get arr in
  let g = Nbk.vec (3 * n) in
  Nbk.grad_v_sum (Nbk.ctx ()) m q !Base.eps2 (0, n, 0, n) g;
  g

let b_operate h bs =
  let f = h /. 6.0 and g = gsum_of (fun b -> b.q) bs in
  List.iteri (fun i { p = p } ->
      for k = 0 to 2 do p.(k) <- p.(k) -. f *. g.{3 * i + k} done) bs

let c_operate h bs =
  let pf = 2.0 *. h /. 3.0 in
  List.iter (fun { q = q; q_star = q_star } -> Array.blit q 0 q_star 0 3) bs;
  let g = gsum_of (fun b -> b.q) bs in
  List.iteri (fun i { m = m; q_star = qs } ->
      let f = (Base.square h) /. (24.0 *. m) in
      for k = 0 to 2 do qs.(k) <- qs.(k) -. f *. g.{3 * i + k} done) bs;
  let g = gsum_of (fun b -> b.q_star) bs in
  List.iteri (fun i { p = p } ->
      for k = 0 to 2 do p.(k) <- p.(k) -. pf *. g.{3 * i + k} done) bs
