"""miniml.stdlib - the part of OCaml's standard library that the reference's hot-path modules and
oracle/ref_fixtures/gen.ml use (TEST INFRASTRUCTURE).  Semantics follow the OCaml manual's
descriptions of Stdlib / Array / List / Printf / Int64 / String; where the real library leaves the
result to the algorithm (ties under an unstable sort) the function refuses instead of guessing.
"""
import functools
import math
import os
import re
import struct

from .interp import (NIL, UNIT, Builtin, Cons, Ctor, MLError, MLExn, Module, Record, apply,
                     compare, equal, fdiv, fpow, from_list, fsqrt, ge, gt, idiv, imod, le, lt, mkref,
                     to_list)

INLINE_OPS = ("+.", "-.", "*.", "/.", "+", "-", "*", "=", "<", ">", "!", "~-.", "~-")
EXTERNALS = {}


def invalid_arg(msg):
    return MLExn(Ctor("Invalid_argument", msg))


def failure(msg):
    return MLExn(Ctor("Failure", msg))


class Channel:
    def __init__(self, fh, name, binary=False):
        self.fh, self.name, self.binary = fh, name, binary

    def write(self, s):
        if self.binary and isinstance(s, str):
            s = s.encode("latin-1")
        elif not self.binary and isinstance(s, bytes):
            s = s.decode("latin-1")
        self.fh.write(s)


# ------------------------------------------------------------------------------------------------
# Printf: "%[flags][width][.prec]conv"; returns a curried function of as many arguments as the
# format has conversions
_fmt_re = re.compile(r"%([-+ #0]*)(\d+|\*)?(?:\.(\d+|\*))?([lLn]?)([a-zA-Z%!@,])")


def parse_format(fmt):
    parts, pos = [], 0
    for m in _fmt_re.finditer(fmt):
        if m.start() > pos:
            parts.append(("lit", fmt[pos:m.start()]))
        pos = m.end()
        flags, width, prec, size, conv = m.groups()
        if conv == "%":
            parts.append(("lit", "%"))
        elif conv == "@":
            parts.append(("lit", "@"))
        elif conv in "!,":
            parts.append(("flush",) if conv == "!" else ("lit", ""))
        else:
            parts.append(("conv", flags or "", width, prec, size, conv))
    if pos < len(fmt):
        parts.append(("lit", fmt[pos:]))
    return parts


def format_one(spec, v):
    _, flags, width, prec, size, conv = spec
    w = width or ""
    p = ("." + prec) if prec is not None else ""
    if conv in "di":
        return ("%" + flags + w + "d") % v
    if conv == "u":
        bits = 64 if size == "L" else (32 if size == "l" else 63)
        return ("%" + flags + w + "d") % (v & ((1 << bits) - 1))
    if conv in "xXo":
        bits = 64 if size == "L" else (32 if size == "l" else 63)
        return ("%" + flags + w + conv) % (v & ((1 << bits) - 1))
    if conv in "feEgG":
        if v != v:
            body = "nan"
        elif v in (math.inf, -math.inf):
            body = "inf" if v > 0 else "-inf"
        else:
            return ("%" + flags + w + p + conv) % v
        return ("%" + flags.replace("0", "") + w + "s") % body
    if conv == "F":
        return float_repr(v)
    if conv == "s":
        return ("%" + flags + w + p + "s") % v
    if conv == "S":
        return '"' + v.replace("\\", "\\\\").replace('"', '\\"') + '"'
    if conv == "c":
        return v
    if conv in "bB":
        return "true" if v else "false"
    raise MLError("Printf: unsupported conversion %%%s" % conv)


def float_repr(x):
    """string_of_float: %.12g, with a trailing '.' when the text looks like an integer."""
    if x != x:
        return "nan"
    if x in (math.inf, -math.inf):
        return "inf" if x > 0 else "-inf"
    s = "%.12g" % x
    return s if any(c in s for c in ".e") else s + "."


def make_printf(sink):
    """sink(text, flush) consumes the finished string; returns printf-like fn(fmt) -> curried."""
    cache = {}

    def printf(fmt):
        parts = cache.get(fmt)
        if parts is None:
            parts = cache[fmt] = parse_format(fmt)
        nargs = 0
        for s in parts:
            if s[0] == "conv":
                nargs += 2 if s[5] == "a" else (1 if s[5] != "t" else 1)

        def finish(*args):
            out, k, flush = [], 0, False
            for s in parts:
                if s[0] == "lit":
                    out.append(s[1])
                elif s[0] == "flush":
                    flush = True
                elif s[5] == "a":
                    raise MLError("Printf: %a is not supported")
                elif s[5] == "t":
                    raise MLError("Printf: %t is not supported")
                else:
                    out.append(format_one(s, args[k]))
                    k += 1
            return sink("".join(out), flush)
        if nargs == 0:
            return finish()
        return Builtin(finish, nargs, "printf-continuation")
    return printf


# ------------------------------------------------------------------------------------------------
# Scanf (sscanf only; the formats the fixture generator uses)
def sscanf(s, fmt, f):
    pos, args, k = 0, [], 0
    while k < len(fmt):
        c = fmt[k]
        if c == "%":
            conv = fmt[k + 1]
            k += 2
            if conv in "di":
                m = re.compile(r"\s*([-+]?\d+)").match(s, pos)
                if not m:
                    raise MLExn(Ctor("Scan_failure", "integer expected"))
                args.append(int(m.group(1)))
            elif conv in "feEgG":
                m = re.compile(r"\s*([-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|inf|nan|infinity))").match(s, pos)
                if not m:
                    raise MLExn(Ctor("Scan_failure", "float expected"))
                args.append(float(m.group(1)))
            elif conv == "s":
                m = re.compile(r"\s*(\S*)").match(s, pos)
                args.append(m.group(1))
            else:
                raise MLError("Scanf: unsupported conversion %%%s" % conv)
            pos = m.end()
        elif c == " ":
            while pos < len(s) and s[pos] in " \t\n\r":
                pos += 1
            k += 1
        else:
            if pos >= len(s) or s[pos] != c:
                raise MLExn(Ctor("Scan_failure", "looking for %r" % c))
            pos += 1
            k += 1
    return apply(f, args)


# ------------------------------------------------------------------------------------------------
# Marshal (output_value): the external format of OCaml's runtime/extern.c, 64-bit, with the
# sharing table keyed on object identity.  Floats stored in records are boxed values in OCaml and
# can be shared there too; which boxes ocamlopt shares depends on its unboxing decisions, so the
# byte stream written here is a VALID encoding of the same value, not necessarily the same bytes.
def marshal_bytes(interp, v, no_sharing=False):
    out = bytearray()
    seen = {}
    state = {"nobj": 0, "size64": 0, "size32": 0}

    def small_or_int(n):
        if 0 <= n < 0x40:
            out.append(0x40 + n)
        elif -128 <= n < 128:
            out.extend(b"\x00" + struct.pack(">b", n))
        elif -32768 <= n < 32768:
            out.extend(b"\x01" + struct.pack(">h", n))
        elif -(1 << 30) <= n < (1 << 30):
            out.extend(b"\x02" + struct.pack(">i", n))
        else:
            out.extend(b"\x03" + struct.pack(">q", n))

    def shared(obj):
        if no_sharing:
            return False
        idx = seen.get(id(obj))
        if idx is None:
            return False
        d = state["nobj"] - idx
        if d < 0x100:
            out.extend(b"\x04" + struct.pack(">B", d))
        elif d < 0x10000:
            out.extend(b"\x05" + struct.pack(">H", d))
        else:
            out.extend(b"\x06" + struct.pack(">I", d))
        return True

    def note(obj, words, words32=None):
        seen[id(obj)] = state["nobj"]
        state["nobj"] += 1
        state["size64"] += 1 + words
        state["size32"] += 1 + (words if words32 is None else words32)

    def block_header(tag, n):
        if tag < 16 and n < 8:
            out.append(0x80 + tag + (n << 4))
        elif n < (1 << 22):
            out.extend(b"\x08" + struct.pack(">I", (n << 10) | tag))
        else:
            out.extend(b"\x13" + struct.pack(">Q", (n << 10) | tag))

    keep = []  # keep every visited object alive so ids stay unique

    def emit(x):
        t = type(x)
        if t is bool:
            small_or_int(1 if x else 0)
        elif t is int:
            small_or_int(x)
        elif x == () and t is tuple:
            small_or_int(0)
        elif t is float:
            if shared(x):
                return
            keep.append(x)
            out.extend(b"\x0c" + struct.pack("<d", x))
            note(x, 1, 2)
        elif t is str:
            if shared(x):
                return
            keep.append(x)
            b = x.encode("latin-1")
            if len(b) < 0x20:
                out.append(0x20 + len(b))
            elif len(b) < 0x100:
                out.extend(b"\x09" + struct.pack(">B", len(b)))
            else:
                out.extend(b"\x0a" + struct.pack(">I", len(b)))
            out.extend(b)
            note(x, len(b) // 8 + 1, len(b) // 4 + 1)
        elif t is list:
            if len(x) == 0:
                out.append(0x80)  # Atom(0)
                return
            if shared(x):
                return
            keep.append(x)
            if all(type(e) is float for e in x):
                if len(x) < 0x100:
                    out.extend(b"\x0e" + struct.pack(">B", len(x)))
                else:
                    out.extend(b"\x07" + struct.pack(">I", len(x)))
                for e in x:
                    out.extend(struct.pack("<d", e))
                note(x, len(x), 2 * len(x))
            else:
                block_header(0, len(x))
                note(x, len(x))
                for e in x:
                    emit(e)
        elif t is Record:
            if shared(x):
                return
            keep.append(x)
            vals = list(x.f.values())
            if vals and all(type(e) is float for e in vals):  # all-float record = flat double array
                out.extend(b"\x0e" + struct.pack(">B", len(vals)))
                for e in vals:
                    out.extend(struct.pack("<d", e))
                note(x, len(vals), 2 * len(vals))
            else:
                block_header(0, len(vals))
                note(x, len(vals))
                for e in vals:
                    emit(e)
        elif t is tuple:
            if shared(x):
                return
            keep.append(x)
            block_header(0, len(x))
            note(x, len(x))
            for e in x:
                emit(e)
        elif t is Cons:
            # a list is a chain of 2-field blocks; written iteratively
            while x is not NIL:
                if shared(x):
                    return
                keep.append(x)
                block_header(0, 2)
                note(x, 2)
                emit(x.hd)
                x = x.tl
            small_or_int(0)
        elif x is NIL:
            small_or_int(0)
        else:
            raise MLError("Marshal: unsupported value %r" % (x,))

    emit(v)
    n = state["nobj"]
    hdr = struct.pack(">5I", 0x8495A6BE, len(out), n, state["size32"], state["size64"])
    return bytes(hdr) + bytes(out)


# ------------------------------------------------------------------------------------------------
def stable_sorted(cmp, xs):
    key = functools.cmp_to_key(lambda a, b: apply(cmp, [a, b]))
    return sorted(xs, key=key)


def check_arr_index(a, ofs, ln, what):
    if ofs < 0 or ln < 0 or ofs + ln > len(a):
        raise invalid_arg(what)


def install(interp):
    root = interp.root

    def val(name, v):
        root.define(name, [v])

    def fn(name, arity, f, module=None):
        b = Builtin(f, arity, name)
        if module is None:
            val(name, b)
        else:
            module.vals[name] = [b]

    def unit(f):
        def g(*a):
            f(*a)
            return UNIT
        return g

    # -- arithmetic
    fn("+", 2, lambda a, b: a + b)
    fn("-", 2, lambda a, b: a - b)
    fn("*", 2, lambda a, b: a * b)
    fn("/", 2, idiv)
    fn("mod", 2, imod)
    fn("land", 2, lambda a, b: a & b)
    fn("lor", 2, lambda a, b: a | b)
    fn("lxor", 2, lambda a, b: a ^ b)
    fn("lsl", 2, lambda a, b: a << b)
    fn("asr", 2, lambda a, b: a >> b)
    fn("+.", 2, lambda a, b: a + b)
    fn("-.", 2, lambda a, b: a - b)
    fn("*.", 2, lambda a, b: a * b)
    fn("/.", 2, fdiv)
    fn("**", 2, fpow)
    fn("~-", 1, lambda a: -a)
    fn("~-.", 1, lambda a: -a)
    fn("~+", 1, lambda a: a)
    fn("~+.", 1, lambda a: a)
    fn("succ", 1, lambda a: a + 1)
    fn("pred", 1, lambda a: a - 1)
    fn("abs", 1, abs)
    fn("abs_float", 1, abs)
    fn("sqrt", 1, fsqrt)

    def guarded(f, name):
        def g(x):
            try:
                return f(x)
            except (ValueError, OverflowError):
                # the C library's answers for the usual suspects
                if name == "exp":
                    return math.inf
                if name in ("log", "log10"):
                    return -math.inf if x == 0.0 else math.nan
                return math.nan
        return g
    for nm in ("exp", "log", "log10", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh",
               "tanh"):
        fn(nm, 1, guarded(getattr(math, nm), nm))
    fn("atan2", 2, math.atan2)
    fn("floor", 1, lambda x: float(math.floor(x)) if math.isfinite(x) else x)
    fn("ceil", 1, lambda x: float(math.ceil(x)) if math.isfinite(x) else x)
    fn("truncate", 1, lambda x: int(x))
    fn("int_of_float", 1, lambda x: int(x))
    fn("float_of_int", 1, float)
    fn("float", 1, float)
    fn("mod_float", 2, math.fmod)
    val("infinity", math.inf)
    val("neg_infinity", -math.inf)
    val("nan", math.nan)
    val("epsilon_float", 2.220446049250313e-16)
    val("max_float", 1.7976931348623157e308)
    val("min_float", 2.2250738585072014e-308)
    val("max_int", (1 << 62) - 1)
    val("min_int", -(1 << 62))

    def classify_float(x):
        if x != x:
            return Ctor("FP_nan")
        if x in (math.inf, -math.inf):
            return Ctor("FP_infinite")
        if x == 0.0:
            return Ctor("FP_zero")
        if abs(x) < 2.2250738585072014e-308:
            return Ctor("FP_subnormal")
        return Ctor("FP_normal")
    fn("classify_float", 1, classify_float)

    # -- comparisons
    fn("=", 2, equal)
    fn("<>", 2, lambda a, b: not equal(a, b))
    fn("<", 2, lt)
    fn(">", 2, gt)
    fn("<=", 2, le)
    fn(">=", 2, ge)
    fn("==", 2, lambda a, b: a is b or (type(a) is int and type(b) is int and a == b))
    fn("!=", 2, lambda a, b: not (a is b or (type(a) is int and type(b) is int and a == b)))
    fn("compare", 2, compare)
    fn("min", 2, lambda a, b: a if le(a, b) else b)
    fn("max", 2, lambda a, b: a if ge(a, b) else b)
    fn("not", 1, lambda a: not a)
    fn("&&", 2, lambda a, b: a and b)
    fn("||", 2, lambda a, b: a or b)

    # -- refs, pairs, misc
    fn("ref", 1, mkref)
    fn("!", 1, lambda r: r.f["contents"])

    def assign(r, v):
        r.f["contents"] = v
        return UNIT
    fn(":=", 2, assign)

    def incr(r):
        r.f["contents"] += 1
        return UNIT

    def decr(r):
        r.f["contents"] -= 1
        return UNIT
    fn("incr", 1, incr)
    fn("decr", 1, decr)
    fn("fst", 1, lambda p: p[0])
    fn("snd", 1, lambda p: p[1])
    fn("ignore", 1, lambda x: UNIT)

    def raise_(e):
        raise MLExn(e)
    fn("raise", 1, raise_)

    def failwith(msg):
        raise failure(msg)
    fn("failwith", 1, failwith)

    def invalid_arg_(msg):
        raise invalid_arg(msg)
    fn("invalid_arg", 1, invalid_arg_)

    def exit_(code):
        raise SystemExit(code)
    fn("exit", 1, exit_)
    fn("^", 2, lambda a, b: a + b)

    def append(a, b):
        xs = to_list(a)
        out = b
        for x in reversed(xs):
            out = Cons(x, out)
        return out
    fn("@", 2, append)
    fn("string_of_int", 1, str)
    fn("string_of_float", 1, float_repr)
    fn("string_of_bool", 1, lambda b: "true" if b else "false")

    def int_of_string(s):
        try:
            return int(s.replace("_", ""), 0)
        except ValueError:
            raise failure("int_of_string")
    fn("int_of_string", 1, int_of_string)

    def float_of_string(s):
        try:
            return float(s.replace("_", ""))
        except ValueError:
            raise failure("float_of_string")
    fn("float_of_string", 1, float_of_string)

    # -- channels
    out_ch = Channel(interp.stdout, "stdout")
    err_ch = Channel(interp.stderr, "stderr")
    val("stdout", out_ch)
    val("stderr", err_ch)
    fn("print_string", 1, unit(lambda s: out_ch.write(s)))
    fn("print_endline", 1, unit(lambda s: out_ch.write(s + "\n")))
    fn("print_newline", 1, unit(lambda u: out_ch.write("\n")))
    fn("print_int", 1, unit(lambda n: out_ch.write(str(n))))
    fn("print_float", 1, unit(lambda x: out_ch.write(float_repr(x))))
    fn("prerr_string", 1, unit(lambda s: err_ch.write(s)))
    fn("prerr_endline", 1, unit(lambda s: err_ch.write(s + "\n")))

    def open_in(path):
        try:
            return Channel(open(path, "r", newline=""), path)
        except OSError as ex:
            raise MLExn(Ctor("Sys_error", "%s: %s" % (path, ex.strerror)))

    def open_out(path, binary=False):
        try:
            return Channel(open(path, "wb" if binary else "w", newline="" if not binary else None),
                           path, binary)
        except OSError as ex:
            raise MLExn(Ctor("Sys_error", "%s: %s" % (path, ex.strerror)))

    def input_line(ch):
        line = ch.fh.readline()
        if line == "":
            raise MLExn(Ctor("End_of_file"))
        return line[:-1] if line.endswith("\n") else line
    fn("open_in", 1, open_in)
    fn("open_out", 1, open_out)
    fn("open_out_bin", 1, lambda p: open_out(p, True))
    fn("input_line", 1, input_line)
    fn("close_in", 1, unit(lambda ch: ch.fh.close()))
    fn("close_out", 1, unit(lambda ch: ch.fh.close()))
    fn("flush", 1, unit(lambda ch: ch.fh.flush()))
    fn("output_string", 2, unit(lambda ch, s: ch.write(s)))

    # -- Array
    A = Module("Array")

    def arr_make(n, x):
        if n < 0:
            raise invalid_arg("Array.make")
        return [x] * n

    def arr_init(n, f):
        if n < 0:
            raise invalid_arg("Array.init")
        return [apply(f, [i]) for i in range(n)]

    def arr_sub(a, ofs, ln):
        check_arr_index(a, ofs, ln, "Array.sub")
        return a[ofs:ofs + ln]

    def arr_blit(src, so, dst, do, ln):
        check_arr_index(src, so, ln, "Array.blit")
        check_arr_index(dst, do, ln, "Array.blit")
        dst[do:do + ln] = src[so:so + ln]
        return UNIT

    def arr_fill(a, ofs, ln, x):
        check_arr_index(a, ofs, ln, "Array.fill")
        for k in range(ofs, ofs + ln):
            a[k] = x
        return UNIT

    def arr_iter(f, a):
        for x in a:
            apply(f, [x])
        return UNIT

    def arr_iteri(f, a):
        for i, x in enumerate(a):
            apply(f, [i, x])
        return UNIT

    def arr_fold_left(f, acc, a):
        for x in a:
            acc = apply(f, [acc, x])
        return acc

    def arr_fold_right(f, a, acc):
        for x in reversed(a):
            acc = apply(f, [x, acc])
        return acc

    def arr_stable_sort(cmp, a):
        a[:] = stable_sorted(cmp, a)
        return UNIT

    def arr_sort(cmp, a):
        # Array.sort is a heap sort: the order of elements that compare equal is the algorithm's.
        # Sort stably, and refuse when a tie would make that visible.
        s = stable_sorted(cmp, a)
        for x, y in zip(s, s[1:]):
            if apply(cmp, [x, y]) == 0 and x is not y and not (type(x) in (int, float) and x == y):
                raise MLError("Array.sort with ties: the heap sort's tie order is not modelled")
        a[:] = s
        return UNIT

    def arr_get(a, i):
        if i < 0 or i >= len(a):
            raise invalid_arg("index out of bounds")
        return a[i]

    def arr_set(a, i, x):
        if i < 0 or i >= len(a):
            raise invalid_arg("index out of bounds")
        a[i] = x
        return UNIT
    fn("length", 1, len, A)
    fn("make", 2, arr_make, A)
    fn("create", 2, arr_make, A)
    fn("init", 2, arr_init, A)
    fn("copy", 1, list, A)
    fn("sub", 3, arr_sub, A)
    fn("blit", 5, arr_blit, A)
    fn("fill", 4, arr_fill, A)
    fn("iter", 2, arr_iter, A)
    fn("iteri", 2, arr_iteri, A)
    fn("map", 2, lambda f, a: [apply(f, [x]) for x in a], A)
    fn("mapi", 2, lambda f, a: [apply(f, [i, x]) for i, x in enumerate(a)], A)
    fn("fold_left", 3, arr_fold_left, A)
    fn("fold_right", 3, arr_fold_right, A)
    fn("to_list", 1, from_list, A)
    fn("of_list", 1, to_list, A)
    fn("append", 2, lambda a, b: a + b, A)
    fn("concat", 1, lambda l: [x for a in to_list(l) for x in a], A)
    fn("make_matrix", 3, lambda n, m, x: [[x] * m for _ in range(n)], A)
    fn("stable_sort", 2, arr_stable_sort, A)
    fn("fast_sort", 2, arr_stable_sort, A)
    fn("sort", 2, arr_sort, A)
    fn("get", 2, arr_get, A)
    fn("set", 3, arr_set, A)

    def same_len(a, b, what):
        if len(a) != len(b):
            raise invalid_arg(what)
    fn("for_all", 2, lambda f, a: all(apply(f, [x]) for x in a), A)
    fn("exists", 2, lambda f, a: any(apply(f, [x]) for x in a), A)
    fn("mem", 2, lambda x, a: any(equal(x, y) for y in a), A)
    fn("for_all2", 3, lambda f, a, b: (same_len(a, b, "Array.for_all2"),
                                       all(apply(f, [x, y]) for x, y in zip(a, b)))[1], A)
    fn("exists2", 3, lambda f, a, b: (same_len(a, b, "Array.exists2"),
                                      any(apply(f, [x, y]) for x, y in zip(a, b)))[1], A)
    fn("iter2", 3, lambda f, a, b: (same_len(a, b, "Array.iter2"),
                                    [apply(f, [x, y]) for x, y in zip(a, b)], UNIT)[2], A)
    fn("map2", 3, lambda f, a, b: (same_len(a, b, "Array.map2"),
                                   [apply(f, [x, y]) for x, y in zip(a, b)])[1], A)
    root.define_mod("Array", A)

    # -- List
    L = Module("List")

    def hd(l):
        if l is NIL:
            raise failure("hd")
        return l.hd

    def tl(l):
        if l is NIL:
            raise failure("tl")
        return l.tl

    def nth(l, n):
        if n < 0:
            raise invalid_arg("List.nth")
        while n > 0 and l is not NIL:
            l, n = l.tl, n - 1
        if l is NIL:
            raise failure("nth")
        return l.hd

    def l_iter(f, l):
        while l is not NIL:
            apply(f, [l.hd])
            l = l.tl
        return UNIT

    def l_iteri(f, l):
        i = 0
        while l is not NIL:
            apply(f, [i, l.hd])
            l, i = l.tl, i + 1
        return UNIT

    def l_fold_left(f, acc, l):
        while l is not NIL:
            acc = apply(f, [acc, l.hd])
            l = l.tl
        return acc

    def l_fold_right(f, l, acc):
        for x in reversed(to_list(l)):
            acc = apply(f, [x, acc])
        return acc

    def l_merge(cmp, l1, l2):
        # the stdlib's rule: take from l1 while cmp h1 h2 <= 0
        a, b, out = to_list(l1), to_list(l2), []
        i = j = 0
        while i < len(a) and j < len(b):
            if apply(cmp, [a[i], b[j]]) <= 0:
                out.append(a[i])
                i += 1
            else:
                out.append(b[j])
                j += 1
        tail = l1
        for _ in range(i):
            tail = tail.tl
        if i >= len(a):
            tail = l2
            for _ in range(j):
                tail = tail.tl
        # share the untouched tail, as the stdlib's recursion does
        for x in reversed(out):
            tail = Cons(x, tail)
        return tail

    def l_find(f, l):
        while l is not NIL:
            if apply(f, [l.hd]):
                return l.hd
            l = l.tl
        raise MLExn(Ctor("Not_found"))

    def l_assoc(k, l):
        while l is not NIL:
            if equal(l.hd[0], k):
                return l.hd[1]
            l = l.tl
        raise MLExn(Ctor("Not_found"))

    def l_partition(f, l):
        yes, no = [], []
        for x in to_list(l):
            (yes if apply(f, [x]) else no).append(x)
        return (from_list(yes), from_list(no))

    def l_map2(f, a, b):
        xs, ys = to_list(a), to_list(b)
        if len(xs) != len(ys):
            raise invalid_arg("List.map2")
        return from_list([apply(f, [x, y]) for x, y in zip(xs, ys)])

    def l_iter2(f, a, b):
        xs, ys = to_list(a), to_list(b)
        if len(xs) != len(ys):
            raise invalid_arg("List.iter2")
        for x, y in zip(xs, ys):
            apply(f, [x, y])
        return UNIT
    fn("length", 1, lambda l: len(to_list(l)), L)
    fn("hd", 1, hd, L)
    fn("tl", 1, tl, L)
    fn("nth", 2, nth, L)
    fn("rev", 1, lambda l: from_list(to_list(l)[::-1]), L)
    fn("append", 2, append, L)
    fn("rev_append", 2, lambda a, b: append(from_list(to_list(a)[::-1]), b), L)
    fn("concat", 1, lambda ll: from_list([x for l in to_list(ll) for x in to_list(l)]), L)
    fn("flatten", 1, lambda ll: from_list([x for l in to_list(ll) for x in to_list(l)]), L)
    fn("iter", 2, l_iter, L)
    fn("iteri", 2, l_iteri, L)
    fn("map", 2, lambda f, l: from_list([apply(f, [x]) for x in to_list(l)]), L)
    fn("mapi", 2, lambda f, l: from_list([apply(f, [i, x]) for i, x in enumerate(to_list(l))]), L)
    fn("rev_map", 2, lambda f, l: from_list([apply(f, [x]) for x in to_list(l)][::-1]), L)
    fn("fold_left", 3, l_fold_left, L)
    fn("fold_right", 3, l_fold_right, L)
    fn("map2", 3, l_map2, L)
    fn("iter2", 3, l_iter2, L)
    fn("for_all", 2, lambda f, l: all(apply(f, [x]) for x in to_list(l)), L)
    fn("exists", 2, lambda f, l: any(apply(f, [x]) for x in to_list(l)), L)
    fn("mem", 2, lambda x, l: any(equal(x, y) for y in to_list(l)), L)
    fn("filter", 2, lambda f, l: from_list([x for x in to_list(l) if apply(f, [x])]), L)
    fn("find_all", 2, lambda f, l: from_list([x for x in to_list(l) if apply(f, [x])]), L)
    fn("find", 2, l_find, L)
    fn("assoc", 2, l_assoc, L)
    fn("partition", 2, l_partition, L)
    fn("split", 1, lambda l: (from_list([p[0] for p in to_list(l)]), from_list([p[1] for p in to_list(l)])), L)
    fn("combine", 2, lambda a, b: from_list(list(zip(to_list(a), to_list(b)))), L)
    fn("sort", 2, lambda cmp, l: from_list(stable_sorted(cmp, to_list(l))), L)  # List.sort IS stable_sort
    fn("stable_sort", 2, lambda cmp, l: from_list(stable_sorted(cmp, to_list(l))), L)
    fn("fast_sort", 2, lambda cmp, l: from_list(stable_sorted(cmp, to_list(l))), L)
    fn("merge", 3, l_merge, L)
    fn("init", 2, lambda n, f: from_list([apply(f, [i]) for i in range(n)]), L)
    root.define_mod("List", L)

    # -- Printf
    P = Module("Printf")

    def to_channel(ch):
        def sink(text, flush):
            ch.write(text)
            if flush:
                ch.fh.flush()
            return UNIT
        return make_printf(sink)
    fn("printf", 1, to_channel(out_ch), P)
    fn("eprintf", 1, to_channel(err_ch), P)
    fn("fprintf", 2, lambda ch, fmt: to_channel(ch)(fmt), P)
    fn("sprintf", 1, make_printf(lambda text, flush: text), P)
    root.define_mod("Printf", P)

    # -- Scanf
    S = Module("Scanf")
    fn("sscanf", 3, sscanf, S)

    def bscanf(*a):
        raise MLError("Scanf.bscanf is not provided (the readers are not on the hot path)")
    fn("bscanf", 3, bscanf, S)
    root.define_mod("Scanf", S)

    # -- String
    St = Module("String")

    def str_sub(s, ofs, ln):
        if ofs < 0 or ln < 0 or ofs + ln > len(s):
            raise invalid_arg("String.sub")
        return s[ofs:ofs + ln]
    fn("length", 1, len, St)
    fn("sub", 3, str_sub, St)
    fn("concat", 2, lambda sep, l: sep.join(to_list(l)), St)
    fn("trim", 1, lambda s: s.strip(" \t\n\r\x0c"), St)
    fn("split_on_char", 2, lambda c, s: from_list(s.split(c)), St)
    fn("make", 2, lambda n, c: c * n, St)
    fn("get", 2, lambda s, i: s[i], St)
    fn("uppercase_ascii", 1, lambda s: s.upper(), St)
    fn("lowercase_ascii", 1, lambda s: s.lower(), St)
    root.define_mod("String", St)

    # -- Int64 (values are Python ints kept in the signed 64-bit range)
    I = Module("Int64")

    def wrap64(n):
        n &= (1 << 64) - 1
        return n - (1 << 64) if n >= (1 << 63) else n

    def i64_of_string(s):
        try:
            n = int(s.replace("_", ""), 0)
        except ValueError:
            raise failure("Int64.of_string")
        if s.lstrip("-+")[:2].lower() in ("0x", "0o", "0b"):
            if n >= (1 << 64):
                raise failure("Int64.of_string")
            return wrap64(n)
        if not -(1 << 63) <= n < (1 << 63):
            raise failure("Int64.of_string")
        return n
    fn("bits_of_float", 1, lambda x: struct.unpack("<q", struct.pack("<d", x))[0], I)
    fn("float_of_bits", 1, lambda n: struct.unpack("<d", struct.pack("<q", n))[0], I)
    fn("of_string", 1, i64_of_string, I)
    fn("to_string", 1, str, I)
    fn("of_int", 1, lambda n: n, I)
    fn("to_int", 1, lambda n: n, I)
    fn("of_float", 1, lambda x: wrap64(int(x)), I)
    fn("to_float", 1, float, I)
    fn("add", 2, lambda a, b: wrap64(a + b), I)
    fn("sub", 2, lambda a, b: wrap64(a - b), I)
    fn("mul", 2, lambda a, b: wrap64(a * b), I)
    fn("logand", 2, lambda a, b: a & b, I)
    fn("logor", 2, lambda a, b: a | b, I)
    fn("logxor", 2, lambda a, b: a ^ b, I)
    fn("shift_left", 2, lambda a, b: wrap64(a << b), I)
    fn("shift_right", 2, lambda a, b: a >> b, I)
    fn("shift_right_logical", 2, lambda a, b: wrap64((a & ((1 << 64) - 1)) >> b), I)
    I.vals["zero"] = [0]
    I.vals["one"] = [1]
    root.define_mod("Int64", I)

    # -- Unix (mkdir only)
    U = Module("Unix")

    def mkdir(path, perm):
        try:
            os.mkdir(path, perm)
        except FileExistsError:
            raise MLExn(Ctor("Unix_error", (Ctor("EEXIST"), "mkdir", path)))
        except OSError as ex:
            raise MLExn(Ctor("Unix_error", (Ctor("EUNKNOWNERR", ex.errno), "mkdir", path)))
        return UNIT
    fn("mkdir", 2, mkdir, U)
    root.define_mod("Unix", U)

    # -- Marshal
    M = Module("Marshal")

    def marshal_to_channel(ch, v, flags):
        no_sharing = any(f.name == "No_sharing" for f in to_list(flags))
        ch.write(marshal_bytes(interp, v, no_sharing))
        return UNIT
    fn("to_channel", 3, marshal_to_channel, M)
    root.define_mod("Marshal", M)

    # -- Sys
    Sy = Module("Sys")
    Sy.vals["argv"] = [["miniml"]]
    Sy.vals["word_size"] = [64]
    root.define_mod("Sys", Sy)

    # -- Hashtbl (keys: ints, floats, strings and tuples of them; one binding per key is kept
    #    visible - `add` shadows like the stdlib's, `remove` pops the newest)
    H = Module("Hashtbl")

    def hkey(k):
        try:
            hash(k)
        except TypeError:
            raise MLError("Hashtbl: only immediate / string / tuple keys are modelled")
        return k

    def h_find(t, k):
        b = t.get(hkey(k))
        if not b:
            raise MLExn(Ctor("Not_found"))
        return b[-1]

    def h_remove(t, k):
        b = t.get(hkey(k))
        if b:
            b.pop()
            if not b:
                del t[k]
        return UNIT
    fn("create", 1, lambda n: {}, H)
    fn("add", 3, unit(lambda t, k, v: t.setdefault(hkey(k), []).append(v)), H)
    fn("replace", 3, unit(lambda t, k, v: t.__setitem__(hkey(k), t.get(k, [None])[:-1] + [v])), H)
    fn("find", 2, h_find, H)
    fn("find_opt", 2, lambda t, k: Ctor("Some", t[k][-1]) if t.get(hkey(k)) else Ctor("None"), H)
    fn("mem", 2, lambda t, k: bool(t.get(hkey(k))), H)
    fn("remove", 2, h_remove, H)
    fn("length", 1, lambda t: sum(len(b) for b in t.values()), H)
    root.define_mod("Hashtbl", H)

    # -- Lazy, Int32, Bigarray.Array1 (as plain arrays: enough to LOAD code that uses them)
    Lz = Module("Lazy")
    fn("force", 1, lambda l: l.force(), Lz)
    root.define_mod("Lazy", Lz)
    I32 = Module("Int32")
    fn("to_int", 1, lambda n: n, I32)
    fn("of_int", 1, lambda n: n, I32)
    root.define_mod("Int32", I32)
    B = Module("Bigarray")
    A1 = Module("Array1")
    for kind, zero in (("float64", 0.0), ("float32", 0.0), ("int32", 0), ("int64", 0), ("int", 0)):
        B.vals[kind] = [Ctor("Kind", zero)]
    B.vals["c_layout"] = [Ctor("C_layout")]
    B.vals["fortran_layout"] = [Ctor("Fortran_layout")]
    fn("create", 3, lambda kind, layout, n: [kind.arg] * n, A1)
    fn("dim", 1, len, A1)
    fn("fill", 2, unit(lambda a, x: a.__setitem__(slice(None), [x] * len(a))), A1)
    fn("get", 2, arr_get, A1)
    fn("set", 3, arr_set, A1)
    fn("of_array", 3, lambda kind, layout, a: list(a), A1)

    def a1_sub(a, ofs, ln):
        raise MLError("Bigarray.Array1.sub shares storage; views are not modelled")
    fn("sub", 3, a1_sub, A1)
    B.mods["Array1"] = A1
    root.define_mod("Bigarray", B)

    # Pervasives / Stdlib name the initial environment itself
    root.mods["Pervasives"] = root.module
    root.mods["Stdlib"] = root.module
