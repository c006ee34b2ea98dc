(* diff.ml - RESTATEMENT of the third-party `Diff` library's DFloat module (test infrastructure).

   farr/direct-nbody's dFloat.ml:20 does `include Diff.DFloat`; that library is not vendored in the
   reference tree and no version of it is pinned anywhere upstream.  What its call sites
   (dFloat_advancer.ml:20-86, dFloat_hermite.ml, dFloat_pushforward.ml:39-76) rely on is restated
   here as the textbook first-order forward-mode dual number, with exactly the formulas and the
   association order of oracle/num.hh (struct Dual):

     value + one tangent;  (a b)' = a' b + a b';  (a / b)' = (a' - (a / b) b') / b;
     (sqrt a)' = a' / (2 sqrt a);  (a ** C c)' = c a^(c-1) a';  comparisons on the value.

   The operator names DFloat exports are taken from how the reference uses them: it writes
   `Pervasives.( * ) 3 i`, `Pervasives.(=) k1 n2`, `Pervasives.(<) id1 id2` wherever it means the
   integer operations, so DFloat shadows + - * / = < > (and dFloat.ml:22-25 then aliases the
   dotted spellings to them); `C x` promotes a float; `jacobian f` differentiates
   f : diff array -> diff array and `lower_jacobian` brings the result back to floats
   (dFloat_pushforward.ml:67).

   With this file in place of the real library, the reference's OWN dFloat_hermite.ml /
   dFloat_pushforward.ml text drives the computation, so the sequencing of the dual-number Hermite
   integrator is pinned; the tangent ARITHMETIC is pinned only against this restatement. *)

module DFloat = struct
  type diff = C of float | D of float * float   (* D (value, tangent) *)

  let value = function C x -> x | D (x, _) -> x
  let tangent = function C _ -> 0.0 | D (_, d) -> d

  let add a b = D (value a +. value b, tangent a +. tangent b)
  let sub a b = D (value a -. value b, tangent a -. tangent b)
  let mul a b = D (value a *. value b, tangent a *. value b +. value a *. tangent b)
  let div a b =
    let q = value a /. value b in
    D (q, (tangent a -. q *. tangent b) /. value b)
  let dsqrt a =
    let s = Pervasives.sqrt (value a) in
    D (s, tangent a /. (2.0 *. s))
  let dpow a b =
    match b with
    | C c -> D (Pervasives.( ** ) (value a) c,
                c *. (Pervasives.( ** ) (value a) (c -. 1.0)) *. tangent a)
    | D _ -> failwith "Diff restatement: a ** b with a non-constant exponent is not used"
  let neg a = D (~-. (value a), ~-. (tangent a))

  (* d f_i / d x_j, one input direction at a time *)
  let jacobian f = fun x ->
    let n = Array.length x in
    let cols = Array.init n (fun j ->
        f (Array.mapi (fun i xi -> if i = j then D (value xi, 1.0) else C (value xi)) x)) in
    let m = Array.length cols.(0) in
    Array.init m (fun i -> Array.init n (fun j -> C (tangent cols.(j).(i))))

  let lower_jacobian jf x =
    Array.map (fun row -> Array.map value row) (jf (Array.map (fun v -> C v) x))

  (* the exported operators (defined last: everything above uses the float ones) *)
  let sqrt = dsqrt
  let ( ** ) = dpow
  let ( ~-. ) = neg
  let ( ~- ) = neg
  let ( = ) a b = Pervasives.( = ) (value a) (value b)
  let ( <> ) a b = Pervasives.( <> ) (value a) (value b)
  let ( < ) a b = Pervasives.( < ) (value a) (value b)
  let ( > ) a b = Pervasives.( > ) (value a) (value b)
  let ( <= ) a b = Pervasives.( <= ) (value a) (value b)
  let ( >= ) a b = Pervasives.( >= ) (value a) (value b)
  let ( + ) = add
  let ( - ) = sub
  let ( * ) = mul
  let ( / ) = div
end
