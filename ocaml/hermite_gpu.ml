(* hermite_gpu.ml - hermite.ml with its pair loops on the GPU; everything else in hermite.ml
   stays as it is.  NOT COMPILED HERE (no OCaml toolchain in the build image).

   [start_bs] (hermite.ml:148-169) is one all-pairs sweep.  [advance] (hermite.ml:180-189) keeps
   the Hermite.b array RESIDENT on the device: it is uploaded once, and
   - up to Nbk.hermite_resident_max () bodies the whole advance_bodies / finish_bs loop
     (hermite.ml:132-143, 171-178) runs in ONE persistent kernel (2.0 - 3.1 us per step);
   - beyond that the scheduler below steps single bodies with one launch per advance_body
     (Nbk.hermite_body_sum: the stepped body's new state rides along with the next call, 48 bytes
     come back) - no re-packing or re-upload of the N bodies per step. *)

open Hermite

let flat bs =
  let arr = Array.of_list bs in
  ( arr,
    Nbk.pack1 (fun b -> b.t) arr, Nbk.pack1 (fun b -> b.m) arr, Nbk.pack3 (fun b -> b.q) arr,
    Nbk.pack3 (fun b -> b.p) arr, Nbk.pack1 (fun b -> b.h) arr, Nbk.pack3 (fun b -> b.f) arr,
    Nbk.pack3 (fun b -> b.df) arr )

let sub3 a i = Array.init 3 (fun k -> a.{3 * i + k})

(* hermite.ml:148-169: f = -gsum, df = -dgsum, then the start time-scale. *)
let start_bs eta bs =
  let arr = Array.of_list bs in
  let n = Array.length arr in
  let m = Nbk.pack1 (fun b -> b.m) arr
  and q = Nbk.pack3 (fun b -> b.q) arr
  and p = Nbk.pack3 (fun b -> b.p) arr in
  let g = Nbk.vec (3 * n) and dg = Nbk.vec (3 * n) in
  Nbk.grad_v_dgrad_v_sum (Nbk.ctx ()) m q p !Base.eps2 (0, n, 0, n) g dg;
  Array.to_list
    (Array.mapi
       (fun i b ->
          let f = Array.init 3 (fun k -> -. g.{3 * i + k})
          and df = Array.init 3 (fun k -> -. dg.{3 * i + k}) in
          let b = { b with f = f; df = df } in
          { b with h = start_timescale eta b })
       arr)

(* advance_body (hermite.ml:95-130) against the resident state: [slot] is b's row on the device,
   [pending] the (row, 13-double state) of the previously stepped body, not yet on the device. *)
let advance_body_resident eta b slot pending =
  let nt = next_time b in
  let qp = predict_position_into (Array.make 3 0.0) b nt in
  let g = Nbk.vec 3 and dg = Nbk.vec 3 in
  let upd, upd_state = match !pending with
    | None -> (-1, Nbk.vec 13)
    | Some (row, st) -> (row, st) in
  Nbk.hermite_body_sum (Nbk.ctx ()) slot nt !Base.eps2 upd upd_state g dg;
  let fp = Array.init 3 (fun k -> -. g.{k}) and dfp = Array.init 3 (fun k -> -. dg.{k}) in
  let { t = t; m = m; q = q; p = p; f = f; df = df; h = h } = b in
  let p_new = Array.make 3 0.0 and q_new = Array.make 3 0.0 in
  for i = 0 to 2 do
    p_new.(i) <- p.(i) +. 0.5 *. h *. (f.(i) +. fp.(i))
                 +. (Base.square h) /. 12.0 *. (df.(i) -. dfp.(i));
    q_new.(i) <- q.(i) +. 0.5 *. h /. m *. (p.(i) +. p_new.(i))
                 +. (Base.square h) /. (12.0 *. m) *. (f.(i) -. fp.(i))
  done;
  incr Base.nsteps;
  let nb = { b with t = t +. h; q = q_new; p = p_new; f = fp; df = dfp;
                    h = timescale eta m q_new qp fp h } in
  let st = Nbk.vec 13 in
  st.{0} <- nb.t;
  for k = 0 to 2 do
    st.{1 + k} <- nb.q.(k); st.{4 + k} <- nb.p.(k); st.{7 + k} <- nb.f.(k); st.{10 + k} <- nb.df.(k)
  done;
  pending := Some (slot, st);
  nb

(* hermite.ml:180-189 *)
let advance bs dt sf =
  let bs = Array.to_list bs in
  let bs = if (List.hd bs).h < 0.0 then start_bs sf bs else bs in
  let arr, t, m, q, p, h, f, df = flat bs in
  let n = Array.length arr in
  let ctx = Nbk.ctx () in
  Nbk.hermite_upload ctx t m q p h f df;                      (* ONE upload *)
  let sorted = List.fast_sort compare_next_time bs in
  let stop_time = (List.hd sorted).t +. dt in
  if n <= Nbk.hermite_resident_max () then begin
    (* the whole stepping loop on the device *)
    let steps = Nbk.hermite_advance_resident ctx stop_time sf !Base.eps2 max_int in
    Base.nsteps := !Base.nsteps + steps;
    Nbk.hermite_download ctx t q p h f df;
    let out = Array.mapi (fun i b ->
        { b with t = t.{i}; h = h.{i}; q = sub3 q i; p = sub3 p i; f = sub3 f i; df = sub3 df i })
        arr in
    Array.fast_sort compare_id out;
    out
  end else begin
    (* host scheduler (hermite.ml:171-178), one launch per advance_body on the resident state *)
    let slot = Hashtbl.create n in
    Array.iteri (fun i b -> Hashtbl.replace slot b.id i) arr;
    let pending = ref None in
    let step b = advance_body_resident sf b (Hashtbl.find slot b.id) pending in
    let rec advance_bodies = function
      | [] -> [||]
      | (b :: rest) as all ->
        if next_time b > stop_time then begin
          (* finish_bs, hermite.ml:132-143 *)
          let rec loop done_bs = function
            | [] -> let a = Array.of_list done_bs in Array.fast_sort compare_id a; a
            | b :: bs -> loop (step { b with h = stop_time -. b.t } :: done_bs) bs in
          loop [] all
        end else
          advance_bodies (List.merge compare_next_time [step b] rest) in
    advance_bodies sorted
  end
