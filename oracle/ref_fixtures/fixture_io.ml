(* fixture_io.ml - bit-pattern I/O shared by gen_wide.ml / gen_ics.ml / gen_dfloat.ml
   (REFERENCE-SIDE test infrastructure; same formats as gen.ml):
   inputs: first line "n <count>", then per body 7 IEEE-754 bit patterns  m qx qy qz px py pz
   outputs: one line per value   <key> <i> <j> <k> <16 hex digits> *)

let bits x = Printf.sprintf "%016Lx" (Int64.bits_of_float x)
let of_bits s = Int64.float_of_bits (Int64.of_string ("0x" ^ s))

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
let emiti oc key i j k n = emit oc key i j k (float_of_int n)
