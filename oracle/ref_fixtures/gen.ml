(* gen.ml - REFERENCE-SIDE fixture generator for oracle/ (test infrastructure, not product code).

   Links the UNMODIFIED modules of farr/direct-nbody (Base, Body, Energy, Hermite, Leapfrog,
   Advancer, Omelyan_advancer), reads the committed seeded inputs (oracle/ref_fixtures/inputs/*.txt:
   IEEE-754 bit patterns, so nothing depends on decimal printing) and writes what the reference's
   own functions return, again as bit patterns, into ref/*.hex:

     ref/pairs.hex    Base.v, Base.grad_v, Base.grad_v_dgrad_v, Leapfrog.kick per pair
                      (base.ml:60-110, leapfrog.ml:70-104)
     ref/sums.hex     Hermite.start_bs (hermite.ml:148-169), Energy.Make(Advancer.A) (energy.ml:57-78)
     ref/traj.hex     Hermite.advance (hermite.ml:180-189), Leapfrog.advance_const_dt / advance
                      (leapfrog.ml:144-165), OA.advance (omelyan_advancer.ml:96-104),
                      Advancer.A.advance (advancer.ml:422-428)
     ref/states_adv.bin   Marshal.to_channel of the fresh and the advanced Advancer.A.body arrays

   The repository cannot run this (its build image has no OCaml toolchain); the first machine that
   has one pins the oracle:   make -C oracle/ref_fixtures REF=/path/to/direct-nbody
   then  python -m pytest tests/test_ref_fixtures.py  compares ref/*.hex with the oracle bit for
   bit (and `diff -r expected ref` shows the same thing without Python).

   Line format:  <key> <i> <j> <k> <16 hex digits>   (unused indices are 0). *)

let bits x = Printf.sprintf "%016Lx" (Int64.bits_of_float x)
let of_bits s = Int64.float_of_bits (Int64.of_string ("0x" ^ s))

(* inputs: first line "n <count>", then per body 7 bit patterns: m qx qy qz px py pz *)
let read_bodies path =
  let ic = open_in path in
  let n = Scanf.sscanf (input_line ic) "n %d" (fun n -> n) in
  let bodies = Array.init n (fun _ ->
      let w = Array.of_list (List.filter (fun s -> s <> "")
                               (String.split_on_char ' ' (String.trim (input_line ic)))) in
      let f k = of_bits w.(k) in
      (f 0, [| f 1; f 2; f 3 |], [| f 4; f 5; f 6 |])) in
  close_in ic;
  bodies

let emit oc key i j k x = Printf.fprintf oc "%s %d %d %d %s\n" key i j k (bits x)
let emit3 oc key i j v = Array.iteri (fun k x -> emit oc key i j k x) v

(* ---- per-pair values ------------------------------------------------------------------------ *)
let pairs bodies =
  let oc = open_out "ref/pairs.hex" in
  let np = 12 in
  List.iteri (fun e eps2 ->
      Base.eps2 := eps2;
      for i = 0 to np - 1 do
        for j = i + 1 to np - 1 do
          let (m1, q1, p1) = bodies.(i) and (m2, q2, p2) = bodies.(j) in
          emit oc (Printf.sprintf "v%d" e) i j 0 (Base.v m1 q1 m2 q2);
          let gv = Array.make 3 0.0 and dgv = Array.make 3 0.0 in
          Base.grad_v m1 q1 m2 q2 gv;
          emit3 oc (Printf.sprintf "grad_v%d" e) i j gv;
          Base.grad_v_dgrad_v m1 q1 p1 m2 q2 p2 gv dgv;
          emit3 oc (Printf.sprintf "gvdgv_g%d" e) i j gv;
          emit3 oc (Printf.sprintf "gvdgv_dg%d" e) i j dgv
        done
      done)
    [0.0; 0.0016];
  Base.eps2 := 0.0;
  (* Leapfrog.kick eta dt on fresh copies with dt_max = infinity (leapfrog.ml ignores eps2) *)
  let eta = 1e-3 and dt = 1e-4 in
  for i = 0 to np - 1 do
    for j = i + 1 to np - 1 do
      let mk (m, q, p) = let b = Leapfrog.make 0.0 m q p in b.Leapfrog.dt_max <- infinity; b in
      let b1 = mk bodies.(i) and b2 = mk bodies.(j) in
      Leapfrog.kick eta dt b1 b2;
      emit3 oc "kick_v1" i j b1.Leapfrog.v;
      emit3 oc "kick_v2" i j b2.Leapfrog.v;
      emit oc "kick_dtmax1" i j 0 b1.Leapfrog.dt_max;
      emit oc "kick_dtmax2" i j 0 b2.Leapfrog.dt_max
    done
  done;
  close_out oc

(* ---- per-body sums ---------------------------------------------------------------------------- *)
module E = Energy.Make (Advancer.A)

let sums bodies =
  let oc = open_out "ref/sums.hex" in
  Base.eps2 := 0.0;
  let hb = List.map (fun (m, q, p) -> Hermite.make 0.0 m q p) (Array.to_list bodies) in
  let started = Hermite.start_bs 1e-2 hb in
  List.iteri (fun i b ->
      emit3 oc "start_f" i 0 b.Hermite.f;
      emit3 oc "start_df" i 0 b.Hermite.df;
      emit oc "start_h" i 0 0 b.Hermite.h) started;
  let ab = Array.map (fun (m, q, p) -> Advancer.A.make 0.0 m q p) bodies in
  emit oc "ke" 0 0 0 (E.total_kinetic_energy ab);
  emit oc "pe" 0 0 0 (E.total_potential_energy ab);
  emit oc "energy" 0 0 0 (E.energy ab);
  close_out oc

(* ---- short trajectories ------------------------------------------------------------------------- *)
let traj bodies =
  let oc = open_out "ref/traj.hex" in
  Base.eps2 := 0.0;
  let n = Array.length bodies in
  (* Hermite.advance bs 0.125 1e-4: output sorted by id (hermite.ml:136) = input order *)
  Base.nsteps := 0;
  let hb = Array.map (fun (m, q, p) -> Hermite.make 0.0 m q p) bodies in
  let out = Hermite.advance hb 0.125 1e-4 in
  Array.iteri (fun i b ->
      emit oc "hermite_t" i 0 0 b.Hermite.t;
      emit oc "hermite_h" i 0 0 b.Hermite.h;
      emit3 oc "hermite_q" i 0 b.Hermite.q;
      emit3 oc "hermite_p" i 0 b.Hermite.p;
      emit3 oc "hermite_f" i 0 b.Hermite.f;
      emit3 oc "hermite_df" i 0 b.Hermite.df) out;
  emit oc "hermite_nsteps" 0 0 0 (float_of_int !Base.nsteps);
  (* Leapfrog: ids are printed relative to the first body made, i.e. as input indices *)
  let lb () = Array.map (fun (m, q, p) -> Leapfrog.make 0.0 m q p) bodies in
  let l0 = lb () in
  let id0 = l0.(0).Leapfrog.id in
  let out = Leapfrog.advance_const_dt l0 (1.0 /. 256.0) 8 in
  Array.iteri (fun s b ->
      emit oc "lfc_id" s 0 0 (float_of_int (b.Leapfrog.id - id0));
      emit oc "lfc_t" s 0 0 b.Leapfrog.t;
      emit oc "lfc_dtmax" s 0 0 b.Leapfrog.dt_max;
      emit3 oc "lfc_q" s 0 b.Leapfrog.q;
      emit3 oc "lfc_v" s 0 b.Leapfrog.v) out;
  let l1 = lb () in
  let id1 = l1.(0).Leapfrog.id in
  let out = Leapfrog.advance l1 (1.0 /. 64.0) 1e-2 in
  Array.iteri (fun s b ->
      emit oc "lfa_id" s 0 0 (float_of_int (b.Leapfrog.id - id1));
      emit oc "lfa_t" s 0 0 b.Leapfrog.t;
      emit oc "lfa_dtmax" s 0 0 b.Leapfrog.dt_max;
      emit3 oc "lfa_q" s 0 b.Leapfrog.q;
      emit3 oc "lfa_v" s 0 b.Leapfrog.v) out;
  (* Omelyan_advancer.OA.advance h *)
  let ob = Array.map (fun (m, q, p) -> Omelyan_advancer.OA.make 0.0 m q p) bodies in
  let out = Omelyan_advancer.OA.advance (1.0 /. 64.0) ob in
  Array.iteri (fun i b ->
      emit oc "oa_t" i 0 0 (Omelyan_advancer.OA.t b);
      emit3 oc "oa_q" i 0 (Omelyan_advancer.OA.q b);
      emit3 oc "oa_p" i 0 (Omelyan_advancer.OA.p b)) out;
  (* Advancer.A.advance bs 0.125 1e-3: output sorted by hmax *)
  let ab = Array.map (fun (m, q, p) -> Advancer.A.make 0.0 m q p) bodies in
  let ida = Advancer.A.id ab.(0) in
  let out = Advancer.A.advance ab 0.125 1e-3 in
  (* the drivers' binary state stream (bin/nbody.ml:80, bin/el_mass_segregation.ml:104-108):
     two values back to back, read here by host/marshal.cpp (tests/test_marshal.py) *)
  let st = open_out_bin "ref/states_adv.bin" in
  Marshal.to_channel st ab [];
  Marshal.to_channel st out [];
  close_out st;
  Array.iteri (fun s b ->
      emit oc "adv_id" s 0 0 (float_of_int (Advancer.A.id b - ida));
      emit oc "adv_t" s 0 0 b.Advancer.A.t;
      emit oc "adv_h" s 0 0 b.Advancer.A.h;
      emit oc "adv_hmax" s 0 0 b.Advancer.A.hmax;
      emit3 oc "adv_q" s 0 b.Advancer.A.q;
      emit3 oc "adv_p" s 0 b.Advancer.A.p) out;
  ignore n;
  close_out oc

let () =
  (try Unix.mkdir "ref" 0o755 with Unix.Unix_error (Unix.EEXIST, _, _) -> ());
  pairs (read_bodies "inputs/plummer_n64_seed20240.txt");
  sums (read_bodies "inputs/plummer_n64_seed20240.txt");
  traj (read_bodies "inputs/plummer_n16_std_seed20241.txt");
  print_endline "wrote ref/pairs.hex ref/sums.hex ref/traj.hex ref/states_adv.bin"
