(* analysis_gpu.ml - the O(N^2) loops of Analysis (analysis.ml:836-876, 904-919) over libnbk.
   binary_energy is evaluated on the device with the reference's IEEE operation sequence, so the
   decisions e < 0 / e < -|eg| and the minima are bit-exact.  NOT COMPILED HERE. *)

open Body

module Make (B : BODY) = struct
  let arrays bs = (Nbk.pack1 B.m bs, Nbk.pack3 B.q bs, Nbk.pack3 B.p bs)

  (* pairs (i, j > i) with binary_energy < threshold, in the reference's loop order *)
  let pairs_below threshold bs =
    let m, q, p = arrays bs in
    let n = Array.length bs in
    let c = Nbk.ctx () in
    (* count first (one scan), then list exactly that many *)
    let emin = Nbk.vec n and partner = Nbk.ivec n and count = Nbk.ivec n in
    Nbk.binary_scan c m q p !Base.eps2 threshold (0, n, 0, n) emin partner count;
    let total = ref 0 in
    for i = 0 to n - 1 do total := !total + Int32.to_int count.{i} done;
    let npairs = !total / 2 in   (* every pair is seen from both ends *)
    let pi = Nbk.ivec (max npairs 1) and pj = Nbk.ivec (max npairs 1)
    and e = Nbk.vec (max npairs 1) in
    let found = Nbk.binary_list c m q p !Base.eps2 threshold pi pj e in
    Array.init found (fun k -> (Int32.to_int pi.{k}, Int32.to_int pj.{k}, e.{k}))

  (* analysis.ml:836-848: a list of ((b1, b2), e_rel), most recent pair first *)
  let binaries bs =
    Array.fold_left (fun acc (i, j, e) -> ((bs.(i), bs.(j)), e) :: acc) [] (pairs_below 0.0 bs)

  (* analysis.ml:850-862 *)
  let binaries_threshold eg bs =
    let l = Array.fold_left (fun acc (i, j, e) -> (bs.(i), bs.(j), e) :: acc) []
        (pairs_below (-. (abs_float eg)) bs) in
    Array.of_list l

  (* analysis.ml:864-876: the most bound pair; ties keep the earlier element of the list *)
  let tightest_binary bs =
    match binaries bs with
    | [] -> raise (Failure "no binaries in system")
    | bin :: bins ->
      fst (List.fold_left
             (fun ((_, e_best) as best) ((_, e) as elt) ->
                if abs_float e > abs_float e_best then elt else best)
             bin bins)

  (* analysis.ml:904-919: [| E_i; E_i/m; |L_i|; |L_i|/m |] per body from ONE potential sweep *)
  let bodies_to_el_phase_space bs =
    let m, q, p = arrays bs in
    let n = Array.length bs in
    let c = Nbk.ctx () in
    Nbk.oa_upload c m q p;
    let el = Nbk.vec (4 * n) in
    ignore (Nbk.oa_diagnostics c !Base.eps2 el);
    Array.init n (fun i -> Array.init 4 (fun k -> el.{4 * i + k}))
end
