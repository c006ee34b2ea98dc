(* gen_wide.ml - REFERENCE-SIDE fixture generator, second sheet (test infrastructure).

   Like gen.ml it links the UNMODIFIED reference modules and writes what their functions return
   as bit patterns; this sheet covers what gen.ml leaves out:
     ref/wide.hex
       eps_*      softened sums and energies: Base.set_eps, Hermite.start_bs (hermite.ml:148-169),
                  Energy.Make over Advancer.A and over Leapfrog (energy.ml:57-78)
       h2_*       two consecutive Hermite.advance calls, softened - the second one resumes from
                  the bodies' own h (hermite.ml:180-189 without start_bs)
       lf3_*      Leapfrog.advance at eta = 1e-3 (deeper block ladder, leapfrog.ml:113-158)
       a2_*       two consecutive Advancer.A.advance calls, softened (advancer.ml:422-428; the
                  second one skips initialize_before_first_step)
       ax_*       Advancer.A.advance ~extpot (advancer.ml:298-314) with the field of the
                  reference's own test (test/advance_test.ml:39-46)
       oa4_*      four OA.advance steps, softened (omelyan_advancer.ml:96-104)
       bin_* thr_* tight_* el_*   Analysis.Make(Advancer.A): binaries, binaries_threshold,
                  tightest_binary, bodies_to_el_phase_space (analysis.ml:836-876, 904-919)
       int_*      Integrator.Make(A)(A).constant_sf_integrate over two checkpoint intervals
                  (integrator.ml:56-84)
   Only stdlib-dependent modules are linked (Analysis uses Random but is not asked to). *)

open Fixture_io

module EA = Energy.Make (Advancer.A)
module EL = Energy.Make (Leapfrog)
module An = Analysis.Make (Advancer.A)
module In = Integrator.Make (Advancer.A) (Advancer.A)

let emit_hermite oc key out =
  Array.iteri (fun i b ->
      emit oc (key ^ "_t") i 0 0 b.Hermite.t;
      emit oc (key ^ "_h") i 0 0 b.Hermite.h;
      emit3 oc (key ^ "_q") i 0 b.Hermite.q;
      emit3 oc (key ^ "_p") i 0 b.Hermite.p;
      emit3 oc (key ^ "_f") i 0 b.Hermite.f;
      emit3 oc (key ^ "_df") i 0 b.Hermite.df) out

let emit_advancer oc key ida out =
  Array.iteri (fun s b ->
      emiti oc (key ^ "_id") s 0 0 (Advancer.A.id b - ida);
      emit oc (key ^ "_t") s 0 0 b.Advancer.A.t;
      emit oc (key ^ "_h") s 0 0 b.Advancer.A.h;
      emit oc (key ^ "_hmax") s 0 0 b.Advancer.A.hmax;
      emit3 oc (key ^ "_q") s 0 b.Advancer.A.q;
      emit3 oc (key ^ "_p") s 0 b.Advancer.A.p) out

let softened oc big small =
  Base.set_eps 0.2;
  let hb = List.map (fun (m, q, p) -> Hermite.make 0.0 m q p) (Array.to_list big) in
  let started = Hermite.start_bs 1e-2 hb in
  List.iteri (fun i b ->
      emit3 oc "eps_start_f" i 0 b.Hermite.f;
      emit3 oc "eps_start_df" i 0 b.Hermite.df;
      emit oc "eps_start_h" i 0 0 b.Hermite.h) started;
  let ab = Array.map (fun (m, q, p) -> Advancer.A.make 0.0 m q p) big in
  emit oc "eps_ke" 0 0 0 (EA.total_kinetic_energy ab);
  emit oc "eps_pe" 0 0 0 (EA.total_potential_energy ab);
  emit oc "eps_energy" 0 0 0 (EA.energy ab);
  let lb = Array.map (fun (m, q, p) -> Leapfrog.make 0.0 m q p) big in
  emit oc "eps_lf_ke" 0 0 0 (EL.total_kinetic_energy lb);
  emit oc "eps_lf_energy" 0 0 0 (EL.energy lb);
  (* two Hermite.advance calls in a row *)
  Base.nsteps := 0;
  let hb = Array.map (fun (m, q, p) -> Hermite.make 0.0 m q p) small in
  let out1 = Hermite.advance hb 0.0625 1e-3 in
  let out2 = Hermite.advance out1 0.0625 1e-3 in
  emit_hermite oc "h2" out2;
  emiti oc "h2_nsteps" 0 0 0 !Base.nsteps;
  (* two Advancer.A.advance calls in a row *)
  let ab = Array.map (fun (m, q, p) -> Advancer.A.make 0.0 m q p) small in
  let ida = Advancer.A.id ab.(0) in
  let out1 = Advancer.A.advance ab 0.0625 1e-3 in
  let out2 = Advancer.A.advance out1 0.0625 1e-3 in
  emit_advancer oc "a2" ida out2;
  emit oc "a2_energy" 0 0 0 (EA.energy out2);
  (* four Omelyan steps *)
  let ob = ref (Array.map (fun (m, q, p) -> Omelyan_advancer.OA.make 0.0 m q p) small) in
  for _step = 1 to 4 do
    ob := Omelyan_advancer.OA.advance (1.0 /. 128.0) !ob
  done;
  Array.iteri (fun i b ->
      emit oc "oa4_t" i 0 0 (Omelyan_advancer.OA.t b);
      emit3 oc "oa4_q" i 0 (Omelyan_advancer.OA.q b);
      emit3 oc "oa4_p" i 0 (Omelyan_advancer.OA.p b)) !ob;
  Base.set_eps 0.0

let unsoftened oc big small =
  Base.set_eps 0.0;
  (* deeper leapfrog ladder *)
  let l = Array.map (fun (m, q, p) -> Leapfrog.make 0.0 m q p) small in
  let id0 = l.(0).Leapfrog.id in
  let out = Leapfrog.advance l (1.0 /. 32.0) 1e-3 in
  Array.iteri (fun s b ->
      emiti oc "lf3_id" s 0 0 (b.Leapfrog.id - id0);
      emit oc "lf3_t" s 0 0 b.Leapfrog.t;
      emit oc "lf3_dtmax" s 0 0 b.Leapfrog.dt_max;
      emit3 oc "lf3_q" s 0 b.Leapfrog.q;
      emit3 oc "lf3_v" s 0 b.Leapfrog.v) out;
  emit oc "lf3_energy" 0 0 0 (EL.energy out);
  (* external potential: the Plummer field of test/advance_test.ml:39-46 (M = 1, eps = 0.1) *)
  let gv b =
    let m = Advancer.A.m b and r = Base.norm (Advancer.A.q b) in
    let r2 = r *. r +. 0.01 in
    let r3 = r2 *. (sqrt r2) in
    Array.map (fun x -> m *. x /. r3) (Advancer.A.q b) in
  let ab = Array.map (fun (m, q, p) -> Advancer.A.make 0.0 m q p) small in
  let ida = Advancer.A.id ab.(0) in
  let out = Advancer.A.advance ~extpot:gv ab 0.0625 1e-2 in
  emit_advancer oc "ax" ida out;
  (* Analysis over the 64-body Plummer sphere *)
  let ab = Array.map (fun (m, q, p) -> Advancer.A.make 0.0 m q p) big in
  let ida = Advancer.A.id ab.(0) in
  List.iteri (fun k ((b1, b2), e) ->
      emiti oc "bin_i" k 0 0 (Advancer.A.id b1 - ida);
      emiti oc "bin_j" k 0 0 (Advancer.A.id b2 - ida);
      emit oc "bin_e" k 0 0 e) (An.binaries ab);
  Array.iteri (fun k (b1, b2, e) ->
      emiti oc "thr_i" k 0 0 (Advancer.A.id b1 - ida);
      emiti oc "thr_j" k 0 0 (Advancer.A.id b2 - ida);
      emit oc "thr_e" k 0 0 e) (An.binaries_threshold 1e-4 ab);
  let (t1, t2) = An.tightest_binary ab in
  emiti oc "tight_i" 0 0 0 (Advancer.A.id t1 - ida);
  emiti oc "tight_j" 0 0 0 (Advancer.A.id t2 - ida);
  Array.iteri (fun i row -> Array.iteri (fun k x -> emit oc "el" i 0 k x) row)
    (An.bodies_to_el_phase_space ab);
  (* Integrator: two checkpoint intervals of 1/32 *)
  let ab = Array.map (fun (m, q, p) -> Advancer.A.make 0.0 m q p) small in
  let ida = Advancer.A.id ab.(0) in
  let out = In.constant_sf_integrate ~ci:0.03125 ~sf:1e-3 ab 0.0625 in
  emit_advancer oc "int" ida out

let () =
  (try Unix.mkdir "ref" 0o755 with Unix.Unix_error (Unix.EEXIST, _, _) -> ());
  let big = read_bodies "inputs/plummer_n64_seed20240.txt"
  and small = read_bodies "inputs/plummer_n16_std_seed20241.txt" in
  let oc = open_out "ref/wide.hex" in
  softened oc big small;
  unsoftened oc big small;
  close_out oc;
  print_endline "wrote ref/wide.hex"
