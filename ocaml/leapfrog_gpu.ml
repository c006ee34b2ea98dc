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

(* ---- Leapfrog.advance with the bodies resident on the device ---------------------------------
   The packed rows (q, m, v) and dt_max live on the device in array order (Nbk.lf_upload); the
   recursion of advance_subsystem (leapfrog.ml:119-142) is walked here while the prefix is wider
   than the device-side limit, and every narrower subtree is ONE call, Nbk.lf_advance_subsystem
   (drifts, kick blocks, dt_max updates and sort_sub on the device; it returns the bodies' new
   times, dt_max and the permutation the sorts applied).  Mirrored by
   direct-nbody_b200/host/leapfrog.cpp (Leap::advance_subsystem_on_device), which tests/ check.
   The wide top of the recursion is left to the per-block path above (kick_block); the host
   mirror shows the same top with nbk_lf_drift / nbk_lf_kick_block / nbk_lf_permute, for which
   stubs are added the same way. *)
let lf_subsystem_max = 1024  (* nbk_lf_subsystem_max () *)

(* advance_subsystem dt eta bs iend for iend <= lf_subsystem_max, bodies already uploaded in
   array order.  Updates t, dt_max and the order of bs.(0..iend-1) in place; q and v stay on the
   device until [fetch]. *)
let advance_subsystem_on_device dt eta bs iend =
  let c = Nbk.ctx () in
  let t = Nbk.vec iend and dtm = Nbk.vec iend in
  let perm = Bigarray.Array1.create Bigarray.int32 Bigarray.c_layout iend in
  for i = 0 to iend - 1 do t.{i} <- bs.(i).t done;
  Nbk.lf_advance_subsystem c dt eta t dtm perm;
  let old = Array.sub bs 0 iend in
  for s = 0 to iend - 1 do
    let b = old.(Int32.to_int perm.{s}) in
    b.t <- t.{s};
    b.dt_max <- dtm.{s};
    bs.(s) <- b
  done

let upload bs =
  let n = Array.length bs in
  Nbk.lf_upload (Nbk.ctx ()) (Nbk.pack1 (fun b -> b.dt_max) bs) (Nbk.pack1 (fun b -> b.m) bs)
    (Nbk.pack3 (fun b -> b.q) bs) (Nbk.pack3 (fun b -> b.v) bs);
  n

let fetch bs =
  let n = Array.length bs in
  let dtm = Nbk.vec n and q = Nbk.vec (3 * n) and v = Nbk.vec (3 * n) in
  Nbk.lf_download (Nbk.ctx ()) dtm q v;
  Array.iteri (fun i b ->
    for k = 0 to 2 do b.q.(k) <- q.{3 * i + k}; b.v.(k) <- v.{3 * i + k} done) bs
