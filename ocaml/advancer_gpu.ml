(* advancer_gpu.ml — Advancer.A (the integrator bin/nbody.ml and bin/el_mass_segregation.ml run)
   over libnbk.  Two levels, both keeping `advance`'s signature and semantics (functional: the
   input array is copied first, advancer.ml:423; result sorted by hmax):

   1. [interact_bodies]: the pair loop of advancer.ml:213-235 as ONE Nbk.interact_sum call; drop
      it in place of A.interact_bodies and nothing else changes (external potentials included).
   2. [advance]: the records live on the device for the whole call.  The top of the block
      recursion (prefixes wider than Nbk.adv_range_max ()) is walked here with one call per step
      phase; every narrower subtree is ONE call, Nbk.adv_advance_range, which runs the whole
      recursion below it on the device.  Without ?extpot only (an external field is an OCaml
      closure; use level 1 for those runs).

   Mirrored, in a language the build image can compile, by direct-nbody_b200/host/advancer.cpp
   (Adv and AdvResident), which tests/ check against the oracle.  NOT COMPILED HERE. *)

open Bigarray
open Advancer.A

let stride = 35

(* ---- level 1 ------------------------------------------------------------------------------- *)
let interact_bodies bs imin imax t vC =
  let n = Array.length bs in
  for i = imin to n - 1 do predict_position bs.(i) t done;
  let sub = Array.sub bs imin (n - imin) in
  let m = Nbk.pack1 (fun b -> b.m) sub and q = Nbk.pack3 (fun b -> b.q) sub in
  let s = Nbk.vec (3 * (n - imin)) in
  Nbk.interact_sum (Nbk.ctx ()) m q (imax - imin) !Base.eps2 s;
  Array.iteri (fun i b ->
    let c0 = vC *. b.cs.(0) and c1 = vC *. b.cs.(1) and c2 = vC *. b.cs.(2) in
    for k = 0 to 2 do
      let sk = s.{3 * i + k} in
      b.dVdq0.(k) <- b.dVdq0.(k) +. c0 *. sk;
      b.dVdq1.(k) <- b.dVdq1.(k) +. c1 *. sk;
      b.dVdq2.(k) <- b.dVdq2.(k) +. c2 *. sk
    done) sub

(* ---- level 2 ------------------------------------------------------------------------------- *)
let vec3 r o v = for k = 0 to 2 do r.{o + k} <- v.(k) done

let to_records bs =
  let r = Nbk.vec (stride * Array.length bs) in
  Array.iteri (fun i b ->
    let o = stride * i in
    r.{o} <- b.t; r.{o + 1} <- b.m; vec3 r (o + 2) b.q; vec3 r (o + 5) b.p;
    r.{o + 8} <- b.h; r.{o + 9} <- b.hmax; r.{o + 10} <- b.t0;
    vec3 r (o + 11) b.dVdq0; vec3 r (o + 14) b.dVdq1; vec3 r (o + 17) b.dVdq2;
    vec3 r (o + 20) b.p0; vec3 r (o + 23) b.q0; vec3 r (o + 26) b.q1; vec3 r (o + 29) b.q2;
    vec3 r (o + 32) b.cs) bs;
  r

let of_record r o b hmax =
  let get3 o = Array.init 3 (fun k -> r.{o + k}) in
  { b with t = r.{o}; m = r.{o + 1}; q = get3 (o + 2); p = get3 (o + 5); h = r.{o + 8}; hmax = hmax;
           t0 = r.{o + 10}; dVdq0 = get3 (o + 11); dVdq1 = get3 (o + 14); dVdq2 = get3 (o + 17);
           p0 = get3 (o + 20); q0 = get3 (o + 23); q1 = get3 (o + 26); q2 = get3 (o + 29);
           cs = get3 (o + 32) }

(* The host keeps what the recursion needs: hmax and the bodies' identity, in row order. *)
type state = { c : Nbk.ctx; hmax : float array; who : int array (* row -> index into bs *) }

let apply_perm st count perm =
  let h = Array.sub st.hmax 0 count and w = Array.sub st.who 0 count in
  for s = 0 to count - 1 do
    st.hmax.(s) <- h.(perm s); st.who.(s) <- w.(perm s)
  done

let compare_hmax h1 h2 = if h1 < h2 then -1 else if h1 > h2 then 1 else 0

(* sort_range bs hmax_lt 0 imax: stable; the device rows follow *)
let sort_prefix st imax =
  let idx = Array.init imax (fun i -> i) in
  Array.stable_sort (fun a b -> compare_hmax st.hmax.(a) st.hmax.(b)) idx;
  if not (Array.for_all2 (=) idx (Array.init imax (fun i -> i))) then begin
    let p = Array1.create int32 c_layout imax in
    Array.iteri (fun s i -> p.{s} <- Int32.of_int i) idx;
    Nbk.adv_permute st.c p;
    apply_perm st imax (fun s -> idx.(s))
  end

(* h_adaptive (advancer.ml:272-281) from a finished record *)
let h_adaptive_of_record sf r o =
  let b = of_record r o (make 0.0 1.0 [|0.;0.;0.|] [|0.;0.;0.|]) nan in
  h_adaptive sf b;
  b.hmax

let rec advance_range st imax dt sf =
  if imax <= Nbk.adv_range_max () then begin
    (* the whole subtree on the device *)
    let h = Nbk.vec imax and p = Array1.create int32 c_layout imax in
    for i = 0 to imax - 1 do h.{i} <- st.hmax.(i) done;
    Nbk.adv_advance_range st.c dt sf !Base.eps2 h p;
    apply_perm st imax (fun s -> Int32.to_int p.{s});
    for i = 0 to imax - 1 do st.hmax.(i) <- h.{i} done
  end else if st.hmax.(imax - 1) < dt then begin
    advance_range st imax (dt /. 2.0) sf;
    advance_range st imax (dt /. 2.0) sf
  end else begin
    let imin =
      let rec loop i = if i < 0 then 0 else if dt < st.hmax.(i) then loop (i - 1) else i + 1 in
      loop (imax - 1) in
    Nbk.adv_begin_step st.c imin imax dt;
    let t0 = Nbk.adv_get_t st.c imin in
    let eps2 = !Base.eps2 in
    Nbk.adv_interact st.c imin imax t0 (dt /. 6.0) eps2;
    if imin > 0 then advance_range st imin (dt /. 2.0) sf;
    Nbk.adv_interact st.c imin imax (t0 +. 0.5 *. dt) (2.0 *. dt /. 3.0) eps2;
    if imin > 0 then advance_range st imin (dt /. 2.0) sf;
    Nbk.adv_interact st.c imin imax (t0 +. dt) (dt /. 6.0) eps2;
    let r = Nbk.vec (stride * (imax - imin)) in
    Nbk.adv_finish st.c imin imax (t0 +. dt) r;
    for k = 0 to imax - imin - 1 do
      st.hmax.(imin + k) <- h_adaptive_of_record sf r (stride * k)
    done;
    sort_prefix st imax
  end

let advance bs dt sf =
  let bs = Array.map copy bs in
  let n = Array.length bs in
  let c = Nbk.ctx () in
  Nbk.adv_upload c (to_records bs);
  let st = { c; hmax = Array.map (fun b -> b.hmax) bs; who = Array.init n (fun i -> i) } in
  if Array.exists first_step bs then begin
    (* initialize_before_first_step (advancer.ml:357-406): the all-pairs sums on the device,
       the first time steps here *)
    Nbk.adv_init_first c !Base.eps2;
    let r = Nbk.vec (stride * n) in
    Nbk.adv_download c 0 n r;
    for i = 0 to n - 1 do
      let b = of_record r (stride * i) bs.(i) nan in
      let fdot = Array.init 3 (fun j -> b.dVdq0.(j) -. b.dVdq1.(j) +. 3.0 *. b.dVdq2.(j)) in
      assign_timestep b (1e-2 *. sf ** 0.33333 *. b.h *. (Base.norm b.dVdq2) /. (Base.norm fdot));
      st.hmax.(i) <- b.hmax
    done
  end;
  sort_prefix st n;
  advance_range st n dt sf;
  let r = Nbk.vec (stride * n) in
  Nbk.adv_download c 0 n r;
  Array.init n (fun s -> of_record r (stride * s) bs.(st.who.(s)) st.hmax.(s))
