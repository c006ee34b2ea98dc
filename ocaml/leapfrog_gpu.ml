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
   The wide top of the recursion uses the per-phase calls on the SAME resident rows
   (Nbk.lf_drift / lf_kick_block / lf_permute): only dt_max comes back per kick block. *)
let lf_subsystem_max = Nbk.lf_subsystem_max ()

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

(* advance_subsystem dt eta bs iend (leapfrog.ml:119-142) on resident bodies, any iend: subtrees
   of up to lf_subsystem_max bodies are one device call; above that the recursion is walked here
   and each phase is one call on the resident rows. *)
let rec advance_subsystem_resident dt eta bs iend =
  if iend <= 0 then ()
  else if iend <= lf_subsystem_max then advance_subsystem_on_device dt eta bs iend
  else begin
    let c = Nbk.ctx () in
    let ibegin = find_can_advance_index bs iend dt eta and dt2 = dt /. 2.0 in
    (* the dt_max reset of the active block (leapfrog.ml:125-127) is done by lf_kick_block *)
    advance_subsystem_resident dt2 eta bs ibegin;
    Nbk.lf_drift c ibegin iend dt2;
    for i = ibegin to iend - 1 do bs.(i).t <- bs.(i).t +. dt2 done;
    let dtm = Nbk.vec iend in
    (* kick loop + the drift that follows it (leapfrog.ml:132-137), applied in place *)
    Nbk.lf_kick_block c iend ibegin eta dt (eta < infinity) dt2 dtm;
    for i = ibegin to iend - 1 do bs.(i).t <- bs.(i).t +. dt2 done;
    for i = 0 to iend - 1 do bs.(i).dt_max <- dtm.{i} done;
    advance_subsystem_resident dt2 eta bs ibegin;
    (* sort_sub (leapfrog.ml:106-111): the device ranks its own dt_max (stable, OCaml's compare)
       and moves the rows; the records here follow the permutation it returns *)
    let perm = Nbk.ivec iend in
    Nbk.lf_sort_sub c perm;
    let old = Array.sub bs 0 iend in
    for s = 0 to iend - 1 do bs.(s) <- old.(Int32.to_int perm.{s}) done
  end

(* Leapfrog.advance (leapfrog.ml:144-158) with resident bodies *)
let advance bs dt eta =
  let bs = Array.map copy bs in
  let n = Array.length bs in
  (* the dt = 0 kick that sets up the time-scales (leapfrog.ml:148-156) *)
  let m = Nbk.pack1 (fun b -> b.m) bs and q = Nbk.pack3 (fun b -> b.q) bs
  and v = Nbk.pack3 (fun b -> b.v) bs in
  let dv = Nbk.vec (3 * n) and tm = Nbk.vec n in
  Nbk.kick_sum (Nbk.ctx ()) m q v eta 0.0 (0, n, 0, n) dv tm;
  Array.iteri (fun i b -> b.dt_max <- tm.{i}) bs;
  Array.fast_sort (fun b1 b2 -> compare b1.dt_max b2.dt_max) bs;
  ignore (upload bs);
  advance_subsystem_resident dt eta bs n;
  fetch bs;
  bs

(* Leapfrog.advance_const_dt (leapfrog.ml:160-165) once every body has dt_max >= dt: all n steps on
   the device, one call (drift dt/2, all-pairs kick, drift dt/2 per step) *)
let advance_const_dt_steady bs dt nsteps =
  let bs = Array.map copy bs in
  let n = Array.length bs in
  let c = Nbk.ctx () in
  Nbk.bodies_upload c (Nbk.pack1 (fun b -> b.m) bs) (Nbk.pack3 (fun b -> b.q) bs)
    (Nbk.pack3 (fun b -> b.v) bs) true;
  let q = Nbk.vec (3 * n) and v = Nbk.vec (3 * n) in
  Nbk.resident_leapfrog_steps c dt nsteps q v;
  Array.iteri (fun i b ->
    b.t <- b.t +. dt *. float_of_int nsteps;
    for k = 0 to 2 do b.q.(k) <- q.{3 * i + k}; b.v.(k) <- v.{3 * i + k} done) bs;
  bs
