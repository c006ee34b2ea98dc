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

(* [advance_steps h n bs]: n calls of OA.advance h (omelyan_advancer.ml:96-104) with the bodies
   resident on the device - one upload, every operator and sweep on the GPU, one download.
   Same bits as n times [advance h]; [f] (if given) sees the total energy after every step
   without the state leaving the device (bin/el_mass_segregation.ml's per-step diagnostics). *)
let advance_steps ?f h nsteps bs =
  let bs = Array.map copy bs in
  let n = Array.length bs and ctx = Nbk.ctx () in
  let m = Nbk.pack1 (fun b -> b.m) bs and q = Nbk.pack3 (fun b -> b.q) bs
  and p = Nbk.pack3 (fun b -> b.p) bs and el = Nbk.vec (4 * n) in
  Nbk.oa_upload ctx m q p;
  for _ = 1 to nsteps do
    Nbk.oa_steps ctx h !Base.eps2 1;
    (match f with
     | Some f -> let ke, pe = Nbk.oa_diagnostics ctx !Base.eps2 el in f (ke +. pe) el
     | None -> ())
  done;
  Nbk.oa_download ctx q p;
  Array.iteri (fun i b ->
      for k = 0 to 2 do b.q.(k) <- q.{3 * i + k}; b.p.(k) <- p.{3 * i + k} done;
      b.t <- b.t +. float_of_int nsteps *. h) bs;
  bs
