"""CPU: oracle/miniml, the interpreter that runs the reference's unmodified OCaml sources here.

The pin of the oracle (tests/test_ref_fixtures.py) is only as good as this interpreter's reading of
OCaml, so the points of the language that decide a floating-point bit or a body's position in an
array are tested one by one against what the OCaml manual specifies: operator precedence and
associativity, unary minus, `**`, integer division and `mod`, IEEE division by zero, polymorphic
compare with nan, stable sorts and List.merge's tie rule, pattern matching, currying, labelled
and optional arguments, records (declared field order, functional update), refs, exceptions,
modules and functors, Printf / Int64 formatting.  Then: every OCaml file this repository ships but
cannot compile (ocaml/*.ml, oracle/ref_fixtures/*.ml) must at least PARSE, and - where the
reference tree is present - LOAD with every name resolved against the reference's modules."""
import glob
import io
import math
import os
import struct
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.miniml import (NIL, Builtin, Interp, MLError, MLExn, ParseError, call, parse_program,  # noqa: E402
                    run_deep, to_list)

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree")


def ev(src, name="v"):
    it = Interp()
    m = it.load_source(src, "T")
    return m.vals[name][0]


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def test_float_precedence_and_association():
    # * / bind tighter than + -, all left-associative; ** tighter still and right-associative
    assert ev("let v = 1.0 +. 2.0 *. 3.0 -. 4.0 /. 8.0") == 1.0 + 2.0 * 3.0 - 4.0 / 8.0
    assert ev("let v = 2.0 ** 3.0 ** 2.0") == 2.0 ** 9.0
    assert ev("let v = 2.0 *. 3.0 ** 2.0") == 18.0
    # association decides the bits: (a+b)+c <> a+(b+c)
    a, b, c = 0.1, 0.2, 0.3
    assert bits(ev("let v = 0.1 +. 0.2 +. 0.3")) == bits((a + b) + c) != bits(a + (b + c))
    assert bits(ev("let v = 0.1 +. (0.2 +. 0.3)")) == bits(a + (b + c))
    x = 1.1
    assert bits(ev("let x = 1.1 let v = x*.x*.x")) == bits((x * x) * x)
    # h*.(e)**0.25 is h *. (e ** 0.25)   (hermite.ml:93)
    assert ev("let h = 3.0 let v = h*.(16.0)**0.25") == 6.0
    # 1e-2*.sf**0.33333*.h : ((1e-2 *. (sf ** 0.33333)) *. h)   (advancer.ml:413)
    sf, h = 1e-3, 0.7
    assert bits(ev("let sf = 1e-3 let h = 0.7 let v = 1e-2*.sf**0.33333*.h")) == bits(
        (1e-2 * math.pow(sf, 0.33333)) * h)
    # a /. b *. c is (a /. b) *. c ; 0.5*.h/.m*.x   (hermite.ml:113)
    assert bits(ev("let v = 0.5*.0.3/.0.7*.1.3")) == bits(((0.5 * 0.3) / 0.7) * 1.3)


def test_arithmetic_grammar_against_an_independent_parser():
    """+ - * / ** and unary minus have the same precedence and associativity in OCaml and in
    Python, so Python's own parser is an independent judge: 400 random expressions, spelled with
    the dotted operators for miniml and the plain ones for `eval`, must agree bit for bit."""
    import random
    rnd = random.Random(12345)

    def atom():
        return repr(round(rnd.uniform(0.1, 9.0), rnd.randint(1, 6)))

    def expr(depth):
        r = rnd.random()
        if depth == 0 or r < 0.2:
            return atom()
        if r < 0.3:
            return "(" + expr(depth - 1) + ")"
        if r < 0.4:
            return "- " + (atom() if rnd.random() < 0.5 else "(" + expr(depth - 1) + ")")
        if r < 0.5:
            return atom() + " ** " + atom()
        return expr(depth - 1) + " " + rnd.choice("+-*/") + " " + expr(depth - 1)
    checked = 0
    while checked < 400:
        e = expr(5)
        try:
            want = eval(e)
        except (ZeroDivisionError, OverflowError):
            continue
        if isinstance(want, complex):
            continue
        ocaml = e.replace("**", "POW")
        for op in "+-*/":
            ocaml = ocaml.replace(" " + op + " ", " " + op + ". ")
        ocaml = ocaml.replace("- ", "-. ").replace("-. . ", "-. ").replace("POW", "**")
        got = ev("let v = " + ocaml)
        assert bits(got) == bits(float(want)), (e, ocaml, got, want)
        checked += 1


def test_unary_minus_and_prefix_operators():
    assert ev("let v = -. 2.0 ** 2.0") == -4.0           # -(2 ** 2)
    assert ev("let x = 3.0 let v = -. x *. 2.0") == -6.0
    assert ev("let m = 2.0 let r = 4.0 let v = ~-.m*.3.0/.r") == -1.5   # base.ml:63
    assert ev("let v = (-1.0)") == -1.0
    assert ev("let f x = x + 1 let v = f 2 - 1") == 2    # application binds tighter than -
    assert ev("let r = ref 5 let v = !r * 2") == 10
    # prefix symbols bind tighter than field / array access: !a.(1) is (!a).(1)
    assert ev("let a = ref [|1; 2|] let v = !a.(1)") == 2
    assert ev("type t = {x : int} let r = ref {x = 7} let v = !r.x") == 7
    assert ev("let v = 7 - -2") == 9


def test_integer_arithmetic_truncates_toward_zero():
    assert ev("let v = (7 / 2, (-7) / 2, 7 / (-2), (-7) mod 3, 7 mod (-3), 1 + 2 * 3)") == (3, -3, -3, -1, 1, 7)
    with pytest.raises(MLExn):
        ev("let v = 1 / 0")


def test_ieee_special_values():
    assert ev("let v = 1.0 /. 0.0") == math.inf and ev("let v = -1.0 /. 0.0") == -math.inf
    assert math.isnan(ev("let v = 0.0 /. 0.0")) and math.isnan(ev("let v = sqrt (-1.0)"))
    assert ev("let v = (nan = nan, nan <> nan, nan < 1.0, compare nan nan, compare nan 0.0)") == (
        False, True, False, 0, -1)
    assert ev("let v = classify_float (1.0 /. 0.0) = FP_infinite && classify_float nan <> FP_normal")
    assert ev("let v = 2.0 ** 0.25") == math.pow(2.0, 0.25)
    assert ev("let v = infinity > max_float && epsilon_float = 2.220446049250313e-16")


def test_comparison_and_boolean_precedence():
    assert ev("let v = 1 + 1 = 2 && 2 < 3 || false") is True
    assert ev("let v = if 1 < 2 then 10 else 20") == 10
    assert ev("let v = not (1 > 2) && 3 >= 3 && 2 <= 3 && 1 <> 2")
    assert ev("let v = (1, 2.0) = (1, 2.0) && [1; 2] = [1; 2] && [|1.0|] <> [|2.0|]")
    assert ev('let v = compare (1, "b") (1, "a")') == 1


def test_if_sequence_and_assignment_precedence():
    # `if c then a <- x; rest` : the ; ends the if
    src = """
    type r = {mutable a : int}
    let x = {a = 0}
    let y = ref 0
    let () = if false then x.a <- 1; y := 5
    let v = (x.a, !y)
    """
    assert ev(src) == (0, 5)
    assert ev("let v = let a = [|0; 0|] in a.(0) <- 1; a.(1) <- a.(0) + 1; a") == [1, 2]
    assert ev("let v = let r = ref 1 in (if !r = 1 then r := 2 else r := 3); !r") == 2
    assert ev("let v = let r = ref 0 in for i = 1 to 4 do r := !r + i done; for i = 3 downto 1 do r := !r * 2 done; !r") == 80
    assert ev("let v = let r = ref 0 in while !r < 10 do incr r done; !r") == 10


def test_currying_closures_and_let_rec():
    assert ev("let add a b c = a + b + c let g = add 1 let h = g 2 let v = h 3") == 6
    assert ev("let twice f x = f (f x) let v = twice (fun x -> x * 3) 2") == 18
    assert ev("let f = fun a -> fun b -> a - b let v = f 5 2") == 3
    assert ev("let rec even n = n = 0 || odd (n - 1) and odd n = n <> 0 && even (n - 1) let v = even 10") is True
    assert ev("let v = let rec fib n = if n < 2 then n else fib (n-1) + fib (n-2) in fib 15") == 610
    # hermite.ml:33-37: a counter closed over by a function
    assert ev("let new_id = let c = ref 0 in fun () -> incr c; !c let a = new_id () let b = new_id () let v = (a, b)") == (1, 2)
    assert ev("let x = 1 let f () = x let x = 2 let v = (f (), x)") == (1, 2)   # lexical scope
    assert ev("let ( +! ) a b = a * 10 + b let v = 1 +! 2") == 12
    assert ev("let v = List.fold_left ( + ) 0 [1; 2; 3]") == 6
    assert ev("let v = (fun (a, b) -> a - b) (5, 3)") == 2


def test_deep_tail_recursion_runs():
    src = "let rec loop n acc = if n = 0 then acc else loop (n - 1) (acc + 1) let v = loop 50000 0"
    assert run_deep(ev, src) == 50000


def test_pattern_matching():
    src = """
    type shape = Dot | Circle of float | Rect of float * float
    let area = function Dot -> 0.0 | Circle r -> r *. r | Rect (w, h) -> w *. h
    let rec len = function [] -> 0 | _ :: t -> 1 + len t
    let classify l = match l with
      | [] -> "empty" | [_] -> "one" | a :: b :: _ when a = b -> "twins" | _ -> "many"
    let first_or d = function Some x -> x | None -> d
    let v = (area Dot, area (Circle 2.0), area (Rect (2.0, 3.0)), len [1;2;3],
             classify [], classify [1], classify [2;2;5], classify [1;2],
             first_or 0 (Some 4), first_or 7 None,
             (match (1, [|2; 3|]) with (a, [|b; c|]) -> a + b + c | _ -> 0),
             (match 3 with 1 | 2 -> "low" | 3 | 4 -> "mid" | _ -> "high"),
             (match "x" with "y" -> 0 | "x" -> 1 | _ -> 2))
    """
    assert ev(src) == (0.0, 4.0, 6.0, 3, "empty", "one", "twins", "many", 4, 7, 6, "mid", 1)
    with pytest.raises(MLExn):
        ev("let v = match 1 with 2 -> 0")


def test_records_and_as_patterns():
    src = """
    type b = {id : int; mutable t : float; q : float array; h : float}
    let mk () = {h = -1.0; q = [|1.0; 2.0|]; id = 7; t = 0.5}
    let next {t = t; h = h} = t +. h
    let bump ({t = t0} as b) = b.t <- t0 +. 1.0; b
    let b = mk ()
    let c = {b with h = 2.0}
    let _ = bump b
    let v = (next b, next c, c.q == b.q, b.id, (let {q = q} = c in q.(1)))
    """
    assert ev(src) == (0.5, 2.5, True, 7, 2.0)   # c was copied before b.t changed; arrays are shared
    it = Interp()
    m = it.load_source(src, "T")
    assert list(m.vals["b"][0].f) == ["id", "t", "q", "h"]   # declared order, not source order


def test_exceptions():
    src = """
    exception Oops of int
    let f x = if x > 2 then raise (Oops x) else x
    let v = (try f 1 with Oops n -> -n), (try f 5 with Oops n -> -n),
            (try failwith "boom" with Failure s -> s),
            (try List.hd [] with Failure _ -> 0),
            (try [|1|].(3) with Invalid_argument _ -> -1),
            (try (try raise Not_found with Failure _ -> 1) with Not_found -> 2)
    """
    assert ev(src) == (1, -5, "boom", 0, -1, 2)
    with pytest.raises(MLExn):
        ev("let v = let a = [|1|] in a.(-1)")


def test_labelled_and_optional_arguments():
    src = """
    let f ?(scale = 2.0) ?shift x = match shift with None -> scale *. x | Some s -> scale *. x +. s
    let g ~a ~b = a - b
    let v = (f 1.0, f ~scale:3.0 1.0, f ~shift:0.5 1.0, f ?shift:(Some 1.0) ~scale:1.0 1.0,
             g ~b:1 ~a:5, (let a = 9 in g ~a ~b:4))
    """
    assert ev(src) == (2.0, 3.0, 2.5, 2.0, 4, 5)


def test_modules_functors_open_include():
    src = """
    module type S = sig type t val v : t -> float val name : string end
    module A = struct type t = float let v x = x *. 2.0 let name = "a" end
    module Make (X : S) : sig val twice : X.t -> float end = struct
      let hidden = 1.0
      let twice x = X.v x +. X.v x +. hidden -. hidden
    end
    module M = Make (A)
    module N = Make (struct type t = float let v x = x +. 1.0 let name = "anon" end)
    module O = struct include A let extra = 5 end
    open O
    let v = (M.twice 1.5, N.twice 1.5, extra, name, A.(v 4.0), Pervasives.( + ) 1 2)
    """
    assert ev(src) == (6.0, 5.0, 5, "a", 8.0, 3)
    # dFloat.ml:22-25: shadowing the float operators for a module's clients
    src = """
    module D = struct
      type d = C of float
      let ( +. ) (C a) (C b) = C (a -. b)
    end
    open D
    let v = (match C 5.0 +. C 3.0 with C x -> x), Pervasives.( +. ) 5.0 3.0
    """
    assert ev(src) == (2.0, 8.0)


def test_sorts_are_stable_and_merge_follows_the_stdlib_tie_rule():
    src = """
    let by_key (a, _) (b, _) = compare a b
    let xs = [(2, "a"); (1, "b"); (2, "c"); (1, "d"); (2, "e")]
    let sorted = List.fast_sort by_key xs
    let arr = Array.of_list xs
    let () = Array.fast_sort by_key arr
    (* List.merge takes from the FIRST list on ties: hermite.ml:178 merges [new_b] in front of
       bodies with the same next time *)
    let merged = List.merge by_key [(1, "new")] [(1, "old"); (2, "z")]
    let merged2 = List.merge by_key [(2, "new")] [(1, "old"); (2, "z"); (3, "y")]
    let v = (sorted, Array.to_list arr, merged, merged2)
    """
    s, a, m1, m2 = ev(src)
    want = [(1, "b"), (1, "d"), (2, "a"), (2, "c"), (2, "e")]
    assert to_list(s) == want and to_list(a) == want
    assert to_list(m1) == [(1, "new"), (1, "old"), (2, "z")]
    assert to_list(m2) == [(1, "old"), (2, "new"), (2, "z"), (3, "y")]
    # compare on floats orders nan first (leapfrog.ml:107 sorts dt_max with compare)
    assert to_list(ev("let v = List.sort compare [2.0; nan; 1.0; infinity]"))[1:] == [1.0, 2.0, math.inf]
    with pytest.raises(MLError):   # heap sort's tie order is not modelled: refuse, do not guess
        ev("let a = [|(1, 1); (1, 2)|] let () = Array.sort (fun (x, _) (y, _) -> compare x y) a let v = a")


def test_list_and_array_library():
    src = """
    let l = [3; 1; 2]
    let v = (List.length l, List.rev l, List.map (fun x -> x * 2) l, l @ [9], 0 :: l,
             List.filter (fun x -> x > 1) l, List.mem 2 l, List.nth l 1,
             Array.init 3 (fun i -> i * i), Array.sub [|1;2;3;4|] 1 2,
             (let a = Array.make 4 0 in Array.blit [|7;8|] 0 a 1 2; Array.fill a 3 1 5; a),
             Array.fold_left (fun acc x -> acc +. x) 0.0 [|0.1; 0.2; 0.3|],
             Array.mapi (fun i x -> float_of_int i *. x) [|1.0; 2.0|],
             List.iteri (fun _ _ -> ()) l, Array.copy [|1|] == [|1|])
    """
    v = ev(src)
    assert v[0] == 3 and to_list(v[1]) == [2, 1, 3] and to_list(v[2]) == [6, 2, 4]
    assert to_list(v[3]) == [3, 1, 2, 9] and to_list(v[4]) == [0, 3, 1, 2] and to_list(v[5]) == [3, 2]
    assert v[6] is True and v[7] == 1 and v[8] == [0, 1, 4] and v[9] == [2, 3] and v[10] == [0, 7, 8, 5]
    assert bits(v[11]) == bits((0.0 + 0.1 + 0.2) + 0.3) and v[12] == [0.0, 2.0] and v[13] == () and v[14] is False
    assert ev("let v = List.tl [1]") is NIL


def test_printf_int64_and_strings():
    src = r"""
    let bits x = Printf.sprintf "%016Lx" (Int64.bits_of_float x)
    let back s = Int64.float_of_bits (Int64.of_string ("0x" ^ s))
    let v = (bits 1.0, bits (-2.5), back "bff0000000000000", bits (back "7ff8000000000001"),
             Printf.sprintf "%s %d %g %g %5.2f|%-4d|%04d %c%%" "k" (-3) 1e-7 100000.0 3.14159 7 42 'z',
             String.split_on_char ' ' (String.trim "  a b  c "), string_of_float 3.0, string_of_int 12,
             Scanf.sscanf "n 17" "n %d" (fun n -> n + 1))
    """
    v = ev(src)
    assert v[0] == "3ff0000000000000" and v[1] == "c004000000000000" and v[2] == -1.0
    assert v[3] == "7ff8000000000001"
    assert v[4] == "k -3 1e-07 100000  3.14|7   |0042 z%"
    assert to_list(v[5]) == ["a", "b", "", "c"] and v[6] == "3." and v[7] == "12" and v[8] == 18
    out = io.StringIO()
    Interp(stdout=out).load_source('let () = Printf.printf "%d-%s\\n" 4 "x"; print_endline "done"', "T")
    assert out.getvalue() == "4-x\ndone\n"


def test_comments_and_lexing():
    assert ev('let v = (* a (* nested *) "*)" comment *) 1 + (* x *) 2') == 3
    assert ev("let a = [|1.0;2.0|] let v = a.(0)-.a.(1)") == -1.0
    assert ev("let r = ref 1 let () = r:=!r+1 let v = !r") == 2
    assert ev("let v = 0x10 + 0o17 + 0b11 + 1_000") == 16 + 15 + 3 + 1000
    assert ev("let v = (1e3, 2., 5e-1)") == (1000.0, 2.0, 0.5)
    assert ev("let f' x = x + 1 let v = f' 1") == 2
    with pytest.raises(ParseError):
        parse_program("let v = (1 + ")
    with pytest.raises(ParseError):   # a `*)` inside a comment ends it - as it does for ocamlc
        parse_program("(* the nbk_oa_*) entry points *) let v = 1")


def test_unprovided_libraries_fail_when_executed_not_when_loaded():
    it = Interp()
    m = it.load_source("let f () = Random.float 1.0 let g () = 2", "T")
    assert call(m.vals["g"][0], ()) == 2
    with pytest.raises(MLError):
        call(m.vals["f"][0], ())


OCAML_FILES = sorted(glob.glob(os.path.join(ROOT, "ocaml", "*.ml")) +
                     glob.glob(os.path.join(ROOT, "oracle", "ref_fixtures", "*.ml")) +
                     glob.glob(os.path.join(ROOT, "oracle", "ref_fixtures", "diff_restated", "*.ml")))


@pytest.mark.parametrize("path", OCAML_FILES, ids=[os.path.relpath(p, ROOT) for p in OCAML_FILES])
def test_shipped_ocaml_files_parse(path):
    """Nothing here can compile OCaml; at least the syntax of every .ml file this repository ships
    is checked (this caught a comment that `*)` closed early in ocaml/nbk.ml)."""
    assert len(OCAML_FILES) >= 13
    assert parse_program(open(path).read(), path)


@needs_ref
def test_ocaml_binding_modules_load_against_the_reference():
    """ocaml/*.ml (the maintainer-side drop-in) load with every identifier resolved against the
    UNMODIFIED reference modules and ocaml/nbk.ml's externals: unbound names, misspelt record
    fields in patterns, wrong module paths and missing arguments to `external`s show up here.
    (Types are not checked - that still needs ocamlc.)"""
    it = Interp(search_path=[REF, os.path.join(ROOT, "ocaml"),
                             os.path.join(ROOT, "oracle", "ref_fixtures", "diff_restated")])
    for prim, val in (("nbk_ml_lf_subsystem_max", 1024), ("nbk_ml_adv_range_max", 1024),
                      ("nbk_ml_hermite_resident_max", 12288)):
        it.externals[prim] = Builtin(lambda u, val=val: val, 1, prim)
    for f in ("nbk", "hermite_gpu", "omelyan_gpu", "leapfrog_gpu", "advancer_gpu", "energy_gpu",
              "analysis_gpu", "dfloat_hermite_gpu"):
        it.load_file(os.path.join(ROOT, "ocaml", f + ".ml"))
    assert not it.missing, it.missing


@needs_ref
def test_every_reference_module_on_the_path_loads():
    it = Interp(search_path=[REF, os.path.join(ROOT, "oracle", "ref_fixtures", "diff_restated")])
    for f in ("base", "body", "energy", "hermite", "leapfrog", "advancer", "omelyan_advancer",
              "integrator", "analysis", "ics", "dFloat", "dFloat_advancer", "dFloat_hermite",
              "dFloat_pushforward"):
        it.load_file(os.path.join(REF, f + ".ml"))
    # what the reference mentions and this interpreter does not provide: a Marshal reader used by
    # Ics.from_file-style loaders and a Diff value used only by the stale half of dFloat_advancer.ml
    assert it.missing <= {"input_value", "zero"}, it.missing


@needs_ref
def test_reference_pair_functions_run_from_their_own_source():
    """Base.grad_v_dgrad_v (base.ml:91-110) executed from the reference's text on a hand-checkable
    pair: two unit masses one length unit apart, approaching at unit speed."""
    it = Interp(search_path=[REF])
    base = it.load_file(os.path.join(REF, "base.ml"))
    gv, dgv = [0.0] * 3, [0.0] * 3
    call(base.vals["grad_v_dgrad_v"][0], 1.0, [0.0, 0.0, 0.0], [0.5, 0.0, 0.0],
         1.0, [1.0, 0.0, 0.0], [-0.5, 0.0, 0.0], gv, dgv)
    # gv = m1 m2 (q1 - q2) / r^3 = (-1, 0, 0);  dgv = dv / r^3 - 3 (r.dv) r / r^5 = (1 - 3, 0, 0)
    assert gv == [-1.0, 0.0, 0.0] and dgv == [-2.0, 0.0, 0.0]
    assert call(base.vals["v"][0], 2.0, [0.0, 0.0, 0.0], 3.0, [0.0, 2.0, 0.0]) == -3.0
