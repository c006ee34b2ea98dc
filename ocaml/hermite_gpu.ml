(* hermite_gpu.ml — the two pair loops of hermite.ml rewritten over Nbk; everything else in
   hermite.ml stays as it is.  Replace the bodies of [start_bs] (hermite.ml:148-169) and the
   List.iter of [advance_body] (hermite.ml:106-115) with these.  NOT COMPILED HERE. *)

open Hermite

(* hermite.ml:148-169: f = -gsum, df = -dgsum, then the start time-scale. *)
let start_bs eta bs =
  let arr = Array.of_list bs in
  let n = Array.length arr in
  let m = Nbk.pack1 (fun b -> b.m) arr
  and q = Nbk.pack3 (fun b -> b.q) arr
  and p = Nbk.pack3 (fun b -> b.p) arr in
  let g = Nbk.vec (3 * n) and dg = Nbk.vec (3 * n) in
  Nbk.grad_v_dgrad_v_sum (Nbk.ctx ()) m q p !Base.eps2 (0, n, 0, n) g dg;
  Array.to_list
    (Array.mapi
       (fun i b ->
          let f = Array.init 3 (fun k -> -. g.{3 * i + k})
          and df = Array.init 3 (fun k -> -. dg.{3 * i + k}) in
          let b = { b with f = f; df = df } in
          { b with h = start_timescale eta b })
       arr)

(* hermite.ml:95-130 with the List.iter (:106-115) replaced: [b] is the body being stepped,
   [bs] the others; fp, dfp come from one 1 x (N-1) predicted sweep. *)
let force_on_predicted b bs =
  let arr = Array.of_list (b :: bs) in
  let n = Array.length arr in
  let nt = next_time b in
  let t = Nbk.pack1 (fun b -> b.t) arr
  and m = Nbk.pack1 (fun b -> b.m) arr
  and q = Nbk.pack3 (fun b -> b.q) arr
  and p = Nbk.pack3 (fun b -> b.p) arr
  and f = Nbk.pack3 (fun b -> b.f) arr
  and df = Nbk.pack3 (fun b -> b.df) arr in
  let g = Nbk.vec 3 and dg = Nbk.vec 3 in
  Nbk.grad_v_dgrad_v_sum_predicted (Nbk.ctx ()) t m q p f df nt !Base.eps2 (0, 1, 0, n) g dg;
  (Array.init 3 (fun k -> -. g.{k}), Array.init 3 (fun k -> -. dg.{k}))
