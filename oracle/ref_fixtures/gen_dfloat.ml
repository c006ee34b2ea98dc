(* gen_dfloat.ml - REFERENCE-SIDE fixture generator for the dual-number Hermite integrator
   (test infrastructure): DFloat_pushforward.Make(DFloat_hermite).jacobian_advance and
   omega_pushforward (dFloat_pushforward.ml:39-76, 140-150) over the UNMODIFIED dFloat.ml,
   dFloat_advancer.ml, dFloat_hermite.ml, dFloat_pushforward.ml.

   Those modules sit on the third-party Diff library, which the reference does not vendor.  Link
   either the real library, or diff_restated/diff.ml (the textbook dual number oracle/num.hh also
   restates): with the latter the reference's own text is pinned, the tangent arithmetic is not. *)

open Fixture_io

module J = DFloat_pushforward.Make (DFloat_hermite)

let run oc key bodies n dt sf =
  let hb = Array.init n (fun i -> let (m, q, p) = bodies.(i) in Hermite.make 0.0 m q p) in
  let jac = J.jacobian_advance dt sf hb in
  Array.iteri (fun i row -> Array.iteri (fun j x -> emit oc (key ^ "_jac") i j 0 x) row) jac;
  emit oc (key ^ "_omega") 0 0 0 (J.omega_pushforward dt sf hb)

let () =
  (try Unix.mkdir "ref" 0o755 with Unix.Unix_error (Unix.EEXIST, _, _) -> ());
  let small = read_bodies "inputs/plummer_n16_std_seed20241.txt" in
  let oc = open_out "ref/dfloat.hex" in
  run oc "n3" small 3 0.03125 1e-2;
  run oc "n5" small 5 0.015625 1e-2;
  close_out oc;
  print_endline "wrote ref/dfloat.hex"
