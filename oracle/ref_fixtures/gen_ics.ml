(* gen_ics.ml - REFERENCE-SIDE fixture generator for the deterministic part of Ics.Make
   (test infrastructure): adjust_frame (ics.ml:161-175) and rescale_to_standard_units
   (ics.ml:304-330) applied to the committed seeded bodies.  A real OCaml build of ics.ml needs
   the gsl package (King models), which is why this sheet is separate from gen_wide.ml. *)

open Fixture_io

module Ic = Ics.Make (Advancer.A)

let emit_bodies oc key bs =
  Array.iteri (fun i b ->
      emit oc (key ^ "_m") i 0 0 (Advancer.A.m b);
      emit3 oc (key ^ "_q") i 0 (Advancer.A.q b);
      emit3 oc (key ^ "_p") i 0 (Advancer.A.p b)) bs

let () =
  (try Unix.mkdir "ref" 0o755 with Unix.Unix_error (Unix.EEXIST, _, _) -> ());
  Base.set_eps 0.0;
  let big = read_bodies "inputs/plummer_n64_seed20240.txt" in
  let ab = Array.map (fun (m, q, p) -> Advancer.A.make 0.0 m q p) big in
  let oc = open_out "ref/ics.hex" in
  let framed = Ic.adjust_frame ab in
  emit_bodies oc "frame" framed;
  emit_bodies oc "std" (Ic.rescale_to_standard_units framed);
  close_out oc;
  print_endline "wrote ref/ics.hex"
