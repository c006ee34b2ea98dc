(* gen_unit.ml - REFERENCE-SIDE fixture generator, third sheet (test infrastructure): energy-error
   TRAJECTORIES over one N-body time unit, the comparison the task's north star asks for
   ("Hermite and leapfrog energy-error trajectories must agree over 1 N-body time unit"), taken
   from the UNMODIFIED Hermite / Leapfrog / Energy modules on a 24-body Plummer sphere in standard
   units: four advances of a quarter time unit each, the total energy (Energy.Make) and the
   state after each.
     ref/unit.hex   he_* Hermite.advance bs 0.25 1e-4 x 4      (hermite.ml:180-189)
                    lf_* Leapfrog.advance bs 0.25 2e-3 x 4     (leapfrog.ml:144-158) *)

open Fixture_io

(* Hermite.b is not a Body.BODY (no id / print / read); energy from its fields directly, with
   Base.v and the fold of energy.ml:57-78 *)
let hermite_energy bs =
  let ke = Array.fold_left (fun e b ->
      e +. (Base.norm_squared b.Hermite.p) /. (2.0 *. b.Hermite.m)) 0.0 bs in
  let pe = ref 0.0 in
  for i = 0 to Array.length bs - 1 do
    for j = i + 1 to Array.length bs - 1 do
      pe := !pe +. Base.v bs.(i).Hermite.m bs.(i).Hermite.q bs.(j).Hermite.m bs.(j).Hermite.q
    done
  done;
  ke +. !pe

module EL = Energy.Make (Leapfrog)

let () =
  (try Unix.mkdir "ref" 0o755 with Unix.Unix_error (Unix.EEXIST, _, _) -> ());
  Base.set_eps 0.0;
  let bodies = read_bodies "inputs/plummer_n24_std_seed20242.txt" in
  let oc = open_out "ref/unit.hex" in
  (* Hermite *)
  Base.nsteps := 0;
  let hb = ref (Array.map (fun (m, q, p) -> Hermite.make 0.0 m q p) bodies) in
  emit oc "he_energy" 0 0 0 (hermite_energy !hb);
  for k = 1 to 4 do
    hb := Hermite.advance !hb 0.25 1e-4;
    emit oc "he_energy" k 0 0 (hermite_energy !hb);
    emiti oc "he_nsteps" k 0 0 !Base.nsteps;
    Array.iteri (fun i b ->
        emit oc "he_t" k i 0 b.Hermite.t;
        emit3 oc "he_q" k i b.Hermite.q;
        emit3 oc "he_p" k i b.Hermite.p) !hb
  done;
  (* Leapfrog *)
  let lb = ref (Array.map (fun (m, q, p) -> Leapfrog.make 0.0 m q p) bodies) in
  let id0 = !lb.(0).Leapfrog.id in
  emit oc "lf_energy" 0 0 0 (EL.energy !lb);
  for k = 1 to 4 do
    lb := Leapfrog.advance !lb 0.25 2e-3;
    emit oc "lf_energy" k 0 0 (EL.energy !lb);
    Array.iteri (fun s b ->
        emiti oc "lf_id" k s 0 (b.Leapfrog.id - id0);
        emit oc "lf_t" k s 0 b.Leapfrog.t;
        emit3 oc "lf_q" k s b.Leapfrog.q;
        emit3 oc "lf_v" k s b.Leapfrog.v) !lb
  done;
  close_out oc;
  print_endline "wrote ref/unit.hex"
