(* leapfrog_gpu.ml — the two kick loops of leapfrog.ml (:132-136 and :151-155) over
   Nbk.kick_sum.  Every pair is evaluated from the velocities at sweep start (SURVEY.md §3.3):
   the velocity kicks agree with the sequential loop to rounding, the dt_max values at O(dt).
   NOT COMPILED HERE. *)

open Leapfrog

let arrays bs iend =
  let sub = Array.sub bs 0 iend in
  (Nbk.pack1 (fun b -> b.m) sub, Nbk.pack3 (fun b -> b.q) sub, Nbk.pack3 (fun b -> b.v) sub)

let apply bs lo hi dv tmin =
  for i = lo to hi - 1 do
    let b = bs.(i) in
    for k = 0 to 2 do b.v.(k) <- b.v.(k) +. dv.{3 * (i - lo) + k} done;
    if b.dt_max > tmin.{i - lo} then b.dt_max <- tmin.{i - lo}
  done

(* for i = ibegin to iend-1 do for j = 0 to i-1 do kick eta dt bs.(i) bs.(j) *)
let kick_block eta dt bs ibegin iend =
  let m, q, v = arrays bs iend in
  let c = Nbk.ctx () in
  let na = iend - ibegin in
  let dv = Nbk.vec (3 * na) and tm = Nbk.vec na in
  Nbk.kick_sum c m q v eta dt (ibegin, iend, 0, iend) dv tm;
  if ibegin > 0 then begin
    let dv' = Nbk.vec (3 * ibegin) and tm' = Nbk.vec ibegin in
    Nbk.kick_sum c m q v eta dt (0, ibegin, ibegin, iend) dv' tm';
    apply bs 0 ibegin dv' tm'
  end;
  apply bs ibegin iend dv tm
