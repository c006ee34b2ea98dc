"""miniml.interp - evaluator for the OCaml subset of miniml.parser (TEST INFRASTRUCTURE).

The parsed program is compiled once into Python closures with statically resolved variables
(one frame per function activation, one cell per module-level binding) and then run.  What
matters for its purpose - executing farr/direct-nbody's unmodified sources so that the CPU
restatement in oracle/ can be compared with what the reference's own text computes - is that

  * `float` is the host's IEEE-754 binary64 and + - * / sqrt are correctly rounded, exactly what
    ocamlopt emits on x86-64 / arm64 (SSE2 / NEON doubles; ocamlopt never contracts a*b+c);
  * the parse tree fixes the association of every floating-point expression (parser.py follows
    the OCaml manual's precedence table);
  * `**` is C's pow(), as in OCaml (Stdlib.( ** ) = caml_power_float = pow);
  * sorts are stable merges with the caller's comparison (Array.stable_sort = Array.fast_sort,
    List.stable_sort, List.merge follow the stdlib's tie rules).

Division by zero, sqrt of negatives etc. produce IEEE results (inf / nan), not Python exceptions.
"""
import math
import sys

from .parser import ParseError, parse_program  # noqa: F401

UNIT = ()


class MLError(Exception):
    """Interpreter-level failure (unbound name, unsupported construct...)."""


class MLExn(Exception):
    """An OCaml exception in flight; .value is the exception value (a Ctor)."""

    def __init__(self, value):
        Exception.__init__(self, value)
        self.value = value

    def __str__(self):
        return "OCaml exception %s" % (show(self.value),)


class Ctor:
    __slots__ = ("name", "arg")

    def __init__(self, name, arg=None):
        self.name, self.arg = name, arg

    def __repr__(self):
        return self.name if self.arg is None else "%s %s" % (self.name, show(self.arg))


class Cons:
    __slots__ = ("hd", "tl")

    def __init__(self, hd, tl):
        self.hd, self.tl = hd, tl


class _Nil:
    __slots__ = ()

    def __repr__(self):
        return "[]"


NIL = _Nil()
NONE = Ctor("None")


class Record:
    __slots__ = ("f",)

    def __init__(self, f):
        self.f = f


def mkref(v):
    return Record({"contents": v})


def from_list(xs):
    out = NIL
    for x in reversed(xs):
        out = Cons(x, out)
    return out


def to_list(l):
    out = []
    while l is not NIL:
        out.append(l.hd)
        l = l.tl
    return out


def show(v):
    if isinstance(v, Record):
        return "{" + "; ".join("%s = %s" % (k, show(x)) for k, x in v.f.items()) + "}"
    if isinstance(v, (Cons, _Nil)):
        return "[" + "; ".join(show(x) for x in to_list(v)) + "]"
    if isinstance(v, list):
        return "[|" + "; ".join(show(x) for x in v) + "|]"
    if isinstance(v, tuple):
        return "(" + ", ".join(show(x) for x in v) + ")"
    return repr(v)


# ------------------------------------------------------------------------------------------------
# functions
class Code:
    __slots__ = ("nslots", "binders", "body", "params", "simple", "npos", "name")


class Closure:
    __slots__ = ("code", "frame")

    def __init__(self, code, frame):
        self.code, self.frame = code, frame


class Builtin:
    __slots__ = ("fn", "arity", "name")

    def __init__(self, fn, arity, name="<builtin>"):
        self.fn, self.arity, self.name = fn, arity, name


class Partial:
    __slots__ = ("f", "args")

    def __init__(self, f, args):
        self.f, self.args = f, args


class LazyVal:
    """lazy e: evaluated at the first Lazy.force, then remembered."""
    __slots__ = ("thunk", "value")

    def __init__(self, thunk):
        self.thunk, self.value = thunk, None

    def force(self):
        if self.thunk is not None:
            self.value = self.thunk()
            self.thunk = None
        return self.value


class Lab:
    """A labelled actual argument: ~name:v (opt=False) or ?name:v (opt=True)."""
    __slots__ = ("name", "value", "opt")

    def __init__(self, name, value, opt):
        self.name, self.value, self.opt = name, value, opt


class LPartial:
    __slots__ = ("clo", "filled")

    def __init__(self, clo, filled):
        self.clo, self.filled = clo, filled


def match_failure(what="pattern"):
    return MLExn(Ctor("Match_failure", what))


def apply(f, args):
    """Apply f to the list of actual arguments (curried: fewer -> partial, more -> re-apply)."""
    while True:
        t = type(f)
        if t is Closure:
            code = f.code
            if not code.simple:
                return apply_labelled(f, {}, args)
            n = code.npos
            k = len(args)
            if k < n:
                return Partial(f, args)
            fr = [None] * code.nslots
            fr[0] = f.frame
            binders = code.binders
            for i in range(n):
                b = binders[i]
                if type(b) is int:
                    fr[b] = args[i]
                elif not b(fr, args[i]):
                    raise match_failure("function parameter")
            if k == n:
                return code.body(fr)
            f = code.body(fr)
            args = args[n:]
        elif t is Builtin:
            n = f.arity
            k = len(args)
            if k < n:
                return Partial(f, args)
            if k == n:
                return f.fn(*args)
            f = f.fn(*args[:n])
            args = args[n:]
        elif t is Partial:
            args = f.args + args
            f = f.f
        elif t is LPartial:
            return apply_labelled(f.clo, f.filled, args)
        else:
            raise MLError("application of a non-function: %s" % show(f))


def apply_labelled(clo, filled, args):
    """Closures with ~label / ?optional parameters.  Optional parameters that were not given are
    defaulted once every positional parameter has its argument (total application)."""
    code = clo.code
    params = code.params
    filled = dict(filled)
    rest = []
    for a in args:
        if type(a) is Lab:
            for i, p in enumerate(params):
                if p[0] != "pos" and p[1] == a.name and i not in filled:
                    if p[0] == "opt":
                        filled[i] = a.value if a.opt else Ctor("Some", a.value)
                    else:
                        filled[i] = a.value
                    break
            else:
                rest.append(a)  # not ours: belongs to the function we return
        else:
            for i, p in enumerate(params):
                if p[0] == "pos" and i not in filled:
                    filled[i] = a
                    break
            else:
                rest.append(a)
    if any(p[0] == "pos" and i not in filled for i, p in enumerate(params)):
        if rest:
            raise MLError("labelled application: surplus arguments before saturation")
        return LPartial(clo, filled)
    fr = [None] * code.nslots
    fr[0] = clo.frame
    for i, p in enumerate(params):
        kind = p[0]
        if i in filled:
            v = filled[i]
        elif kind == "opt":
            v = NONE
        else:
            raise MLError("missing labelled argument ~%s" % p[1])
        if kind == "opt" and p[3] is not None:  # ?(x = default)
            v = v.arg if v.name == "Some" else p[3](fr)
        b = code.binders[i]
        if type(b) is int:
            fr[b] = v
        elif not b(fr, v):
            raise match_failure("function parameter")
    res = code.body(fr)
    return apply(res, rest) if rest else res


def call(f, *args):
    return apply(f, list(args))


# ------------------------------------------------------------------------------------------------
# polymorphic comparison (Stdlib.compare): floats order with nan below everything and equal to
# itself; blocks compare lexicographically
def compare(a, b):
    ta = type(a)
    if ta is float or ta is int or ta is bool:
        if a < b:
            return -1
        if a > b:
            return 1
        if a == b:
            return 0
        # a nan is involved
        an, bn = a != a, b != b
        return 0 if (an and bn) else (-1 if an else 1)
    if ta is str:
        return -1 if a < b else (1 if a > b else 0)
    if ta is tuple or ta is list:
        if ta is list and len(a) != len(b):
            return -1 if len(a) < len(b) else 1
        for x, y in zip(a, b):
            c = compare(x, y)
            if c:
                return c
        return 0
    if ta is Record:
        for x, y in zip(a.f.values(), b.f.values()):
            c = compare(x, y)
            if c:
                return c
        return 0
    if ta is Ctor:
        if type(b) is not Ctor:
            raise MLError("compare: incomparable values")
        # constant constructors sort before non-constant ones; otherwise declaration order, which
        # callers that rely on it must register (ctor_order)
        if a.name != b.name:
            oa, ob = ctor_order.get(a.name), ctor_order.get(b.name)
            if oa is None or ob is None:
                raise MLError("compare: constructor order of %s / %s unknown" % (a.name, b.name))
            return -1 if oa < ob else 1
        if a.arg is None:
            return 0
        return compare(a.arg, b.arg)
    if ta is Cons or ta is _Nil:
        while True:
            if a is NIL:
                return 0 if b is NIL else -1
            if b is NIL:
                return 1
            c = compare(a.hd, b.hd)
            if c:
                return c
            a, b = a.tl, b.tl
    if ta in (Closure, Builtin, Partial, LPartial):
        raise MLExn(Ctor("Invalid_argument", "compare: functional value"))
    raise MLError("compare: unsupported value %r" % (a,))


ctor_order = {}  # constructor name -> (is_nonconstant, index) from the type declarations seen


def equal(a, b):
    """Structural equality: like compare = 0 except that nan <> nan."""
    ta = type(a)
    if ta is float or ta is int or ta is bool or ta is str:
        return a == b
    if ta is tuple or ta is list:
        if len(a) != len(b):
            return False
        for x, y in zip(a, b):
            if not equal(x, y):
                return False
        return True
    if ta is Record:
        return all(equal(x, y) for x, y in zip(a.f.values(), b.f.values()))
    if ta is Ctor:
        if a.name != b.name:
            return False
        return True if a.arg is None else equal(a.arg, b.arg)
    if ta is Cons or ta is _Nil:
        while True:
            if a is NIL or b is NIL:
                return a is b
            if not equal(a.hd, b.hd):
                return False
            a, b = a.tl, b.tl
    if ta in (Closure, Builtin, Partial, LPartial):
        raise MLExn(Ctor("Invalid_argument", "compare: functional value"))
    raise MLError("equal: unsupported value %r" % (a,))


def _num(a):
    t = type(a)
    return t is float or t is int


def lt(a, b):
    return a < b if _num(a) and _num(b) else compare_ord(a, b) < 0


def le(a, b):
    return a <= b if _num(a) and _num(b) else compare_ord(a, b) <= 0


def gt(a, b):
    return a > b if _num(a) and _num(b) else compare_ord(a, b) > 0


def ge(a, b):
    return a >= b if _num(a) and _num(b) else compare_ord(a, b) >= 0


def compare_ord(a, b):
    """The ordering used by < <= > >= on structured values.  (They differ from `compare` only
    when a nan is met, where every comparison is false; structured values holding nans are not
    compared by the sources this interpreter serves, so the distinction is not modelled.)"""
    return compare(a, b)


# IEEE arithmetic where Python raises instead
def fdiv(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return math.nan
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf


def fpow(a, b):
    try:
        return math.pow(a, b)
    except OverflowError:
        # |result| overflows: sign as pow() gives it
        if a < 0 and float(b).is_integer() and int(b) % 2 == 1:
            return -math.inf
        return math.inf
    except ValueError:
        # domain error: negative base with a non-integer exponent -> nan; 0 ** negative -> inf
        if a == 0.0:
            if float(b).is_integer() and int(b) % 2 == 1:
                return math.copysign(math.inf, a)
            return math.inf
        return math.nan


def fsqrt(a):
    try:
        return math.sqrt(a)
    except ValueError:
        return math.nan


def idiv(a, b):
    if b == 0:
        raise MLExn(Ctor("Division_by_zero"))
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


def imod(a, b):
    if b == 0:
        raise MLExn(Ctor("Division_by_zero"))
    r = abs(a) % abs(b)
    return -r if a < 0 else r


# ------------------------------------------------------------------------------------------------
# modules and scopes
class Module:
    def __init__(self, name):
        self.name = name
        self.vals = {}   # name -> cell ([value])
        self.mods = {}   # name -> Module | Functor


class Functor:
    def __init__(self, params, body, scope, interp):
        self.params, self.body, self.scope, self.interp = params, body, scope, interp


class ModScope:
    """Compile-time view of the module being defined: everything visible (own definitions, opened
    modules, the enclosing structure through `parent`) and what it exports (`module`)."""

    def __init__(self, parent, module):
        self.parent = parent
        self.module = module
        self.vals = {}
        self.mods = {}

    def find_val(self, name):
        s = self
        while s is not None:
            c = s.vals.get(name)
            if c is not None:
                return c
            s = s.parent
        return None

    def find_mod(self, name):
        s = self
        while s is not None:
            m = s.mods.get(name)
            if m is not None:
                return m
            s = s.parent
        return None

    def define(self, name, cell):
        self.vals[name] = cell
        self.module.vals[name] = cell

    def define_mod(self, name, m):
        self.mods[name] = m
        self.module.mods[name] = m

    def open(self, m):
        self.vals.update(m.vals)
        self.mods.update(m.mods)


class FnScope:
    """Slots of one function activation.  Slot 0 is the link to the defining activation."""

    def __init__(self, parent):
        self.parent = parent
        self.names = {}
        self.nslots = 1

    def new_slot(self, name):
        i = self.nslots
        self.nslots += 1
        self.names[name] = i
        return i


UNSET = object()


class Interp:
    def __init__(self, search_path=(), stdout=None, stderr=None):
        from . import stdlib
        self.search_path = list(search_path)
        self.stdout = stdout if stdout is not None else sys.stdout
        self.stderr = stderr if stderr is not None else sys.stderr
        self.record_types = {}   # field name -> ordered field list of the LAST record type declaring it
        self.root_module = Module("<root>")
        self.root = ModScope(None, self.root_module)
        self.loading = set()
        self.missing = set()     # names that did not resolve (their uses fail when executed)
        self.externals = {}      # C primitive name -> Builtin standing in for it
        stdlib.install(self)
        # the operators' pristine cells, so that compiled code can inline them when not shadowed
        self.prim_cells = {op: self.root.vals[op] for op in stdlib.INLINE_OPS}

    # -- loading ---------------------------------------------------------------------------------
    def load_file(self, path, modname=None):
        import os
        if modname is None:
            base = os.path.basename(path)
            modname = base[0].upper() + base[1:].rsplit(".", 1)[0]
        with open(path) as fh:
            src = fh.read()
        return self.load_source(src, modname, path)

    def load_source(self, src, modname, fname="<src>"):
        items = parse_program(src, fname)
        m = Module(modname)
        scope = ModScope(self.root, m)
        self.run_items(items, scope)
        self.root.define_mod(modname, m)
        return m

    def find_module_file(self, name):
        import os
        for d in self.search_path:
            for cand in (name[0].lower() + name[1:] + ".ml", name + ".ml"):
                p = os.path.join(d, cand)
                if os.path.exists(p):
                    return p
        return None

    def resolve_mod(self, scope, path):
        m = scope.find_mod(path[0])
        if m is None:
            # a compilation unit not loaded yet: look for <name>.ml on the search path
            p = self.find_module_file(path[0])
            if p is None or path[0] in self.loading:
                raise MLError("unbound module %s" % path[0])
            self.loading.add(path[0])
            try:
                m = self.load_file(p, path[0])
            finally:
                self.loading.discard(path[0])
        for name in path[1:]:
            sub = m.mods.get(name)
            if sub is None:
                raise MLError("unbound module %s in %s" % (name, ".".join(path)))
            m = sub
        return m

    # -- structure items ---------------------------------------------------------------------------
    def run_items(self, items, scope):
        for it in items:
            kind = it[0]
            if kind == "let":
                self.run_let(it[1], it[2], scope)
            elif kind == "eval":
                fs = FnScope(None)
                c = self.compile(it[1], fs, scope)
                c([None] * fs.nslots)
            elif kind == "type":
                for d in it[1]:
                    if d[0] == "record":
                        names = [f for f, _ in d[2]]
                        for f in names:
                            self.record_types[f] = names
                    elif d[0] == "variant":
                        nconst = nblock = 0
                        for cname, has_arg in d[2]:
                            if has_arg:
                                ctor_order[cname] = (1, nblock)
                                nblock += 1
                            else:
                                ctor_order[cname] = (0, nconst)
                                nconst += 1
            elif kind in ("exception", "modtype"):
                pass
            elif kind == "open":
                scope.open(self.resolve_mod(scope, it[1]))
            elif kind == "include":
                m = self.eval_modexpr(it[1], scope)
                for k, c in m.vals.items():
                    scope.define(k, c)
                for k, c in m.mods.items():
                    scope.define_mod(k, c)
            elif kind == "module":
                m = self.eval_modexpr(it[2], scope, it[1])
                scope.define_mod(it[1], m)
            elif kind == "external":
                from . import stdlib
                prim = None
                for pname in it[2]:
                    prim = prim or self.externals.get(pname) or stdlib.EXTERNALS.get(pname)
                if prim is None:
                    # C code this interpreter cannot run: the binding loads, a call fails
                    msg = "external %s = %s is C code; no Python stand-in registered" % (
                        it[1], " ".join(repr(x) for x in it[2]))

                    def unavailable(_x, msg=msg):
                        raise MLError(msg)
                    prim = Builtin(unavailable, 1, it[1])
                scope.define(it[1], [prim])
            else:
                raise MLError("unsupported structure item %s" % kind)

    def eval_modexpr(self, me, scope, name="<anon>"):
        kind = me[0]
        if kind == "struct":
            m = Module(name)
            self.run_items(me[1], ModScope(scope, m))
            return m
        if kind == "mpath":
            return self.resolve_mod(scope, me[1])
        if kind == "functor":
            return Functor(me[1], me[2], scope, self)
        if kind == "mapply":
            f = self.eval_modexpr(me[1], scope)
            arg = self.eval_modexpr(me[2], scope)
            if not isinstance(f, Functor):
                raise MLError("functor application of a non-functor")
            inner = ModScope(f.scope, Module("<functor-arg>"))
            inner.mods[f.params[0]] = arg
            if len(f.params) > 1:
                return Functor(f.params[1:], f.body, inner, self)
            return self.eval_modexpr(f.body, inner, name)
        raise MLError("unsupported module expression %s" % kind)

    def run_let(self, rec, binds, scope):
        if rec:
            cells = []
            for pat, _ in binds:
                if pat[0] != "pvar":
                    raise MLError("let rec needs plain names")
                cell = [UNSET]
                cells.append(cell)
                scope.define(pat[1], cell)
            for cell, (pat, e) in zip(cells, binds):
                fs = FnScope(None)
                c = self.compile(e, fs, scope, name=pat[1])
                cell[0] = c([None] * fs.nslots)
            return
        results = []
        for pat, e in binds:
            fs = FnScope(None)
            c = self.compile(e, fs, scope, name=pat[1] if pat[0] == "pvar" else None)
            v = c([None] * fs.nslots)
            if pat[0] == "pvar":
                results.append([(pat[1], v)])
                continue
            ps = FnScope(None)
            pm = self.compile_pat(pat, ps)
            fr = [None] * ps.nslots
            if not pm(fr, v):
                raise match_failure("top-level let")
            results.append([(n, fr[i]) for n, i in ps.names.items()])
        for group in results:
            for n, v in group:
                scope.define(n, [v])

    # -- patterns ------------------------------------------------------------------------------------
    def compile_pat(self, p, fs):
        """-> matcher(fr, v) -> bool, binding the pattern's variables into slots of fs; a plain
        variable compiles to its slot number (an int) so callers can store directly."""
        kind = p[0]
        if kind == "pvar":
            return fs.new_slot(p[1])
        m = self.compile_pat_fn(p, fs)
        return m

    def compile_pat_fn(self, p, fs):
        kind = p[0]
        if kind == "pvar":
            i = fs.new_slot(p[1])

            def m_var(fr, v):
                fr[i] = v
                return True
            return m_var
        if kind == "pany":
            return lambda fr, v: True
        if kind == "punit":
            return lambda fr, v: True
        if kind == "pconst":
            c = p[1]
            if type(c) is float:
                return lambda fr, v: v == c
            return lambda fr, v: v == c and type(v) is type(c)
        if kind == "ptuple":
            subs = [self.compile_pat_fn(q, fs) for q in p[1]]
            n = len(subs)

            def m_tuple(fr, v):
                if len(v) != n:
                    raise MLError("tuple pattern arity")
                for s, x in zip(subs, v):
                    if not s(fr, x):
                        return False
                return True
            return m_tuple
        if kind == "pctor":
            name = p[1]
            if p[2] is None:
                return lambda fr, v: type(v) is Ctor and v.name == name
            sub = self.compile_pat_fn(p[2], fs)
            return lambda fr, v: type(v) is Ctor and v.name == name and v.arg is not None and sub(fr, v.arg)
        if kind == "pnil":
            return lambda fr, v: v is NIL
        if kind == "pcons":
            a = self.compile_pat_fn(p[1], fs)
            b = self.compile_pat_fn(p[2], fs)
            return lambda fr, v: v is not NIL and a(fr, v.hd) and b(fr, v.tl)
        if kind == "parray":
            subs = [self.compile_pat_fn(q, fs) for q in p[1]]
            n = len(subs)
            return lambda fr, v: len(v) == n and all(s(fr, x) for s, x in zip(subs, v))
        if kind == "precord":
            subs = [(f, self.compile_pat_fn(q, fs)) for f, q in p[1]]

            def m_record(fr, v):
                d = v.f
                for f, s in subs:
                    if not s(fr, d[f]):
                        return False
                return True
            return m_record
        if kind == "pas":
            sub = self.compile_pat_fn(p[1], fs)
            i = fs.new_slot(p[2])

            def m_as(fr, v):
                fr[i] = v
                return sub(fr, v)
            return m_as
        if kind == "por":
            # both sides bind the same names: compile the right side against the left's slots
            a = self.compile_pat_fn(p[1], fs)
            saved = dict(fs.names)
            b = self.compile_pat_fn(p[2], fs)
            right = {n: i for n, i in fs.names.items() if saved.get(n) != i}
            moves = [(right[n], saved[n]) for n in right if n in saved]
            fs.names.update({n: saved[n] for n in right if n in saved})

            def m_or(fr, v):
                if a(fr, v):
                    return True
                if b(fr, v):
                    for src, dst in moves:
                        fr[dst] = fr[src]
                    return True
                return False
            return m_or
        raise MLError("unsupported pattern %s" % kind)

    # -- expressions ----------------------------------------------------------------------------------
    def resolve_var(self, name, fs, scope):
        """-> ('local', depth, slot) | ('cell', cell)"""
        depth, s = 0, fs
        while s is not None:
            i = s.names.get(name)
            if i is not None:
                return ("local", depth, i)
            s = s.parent
            depth += 1
        cell = scope.find_val(name)
        if cell is None:
            raise MLError("unbound value %s" % name)
        return ("cell", cell)

    def compile_var(self, name, fs, scope):
        try:
            r = self.resolve_var(name, fs, scope)
        except MLError as ex:
            # a standard-library value this interpreter does not provide: fail when executed
            msg = str(ex)
            self.missing.add(name)

            def missing(fr):
                raise MLError(msg)
            return missing
        if r[0] == "local":
            _, depth, i = r
            if depth == 0:
                return lambda fr: fr[i]
            if depth == 1:
                return lambda fr: fr[0][i]
            if depth == 2:
                return lambda fr: fr[0][0][i]

            def deep(fr):
                for _ in range(depth):
                    fr = fr[0]
                return fr[i]
            return deep
        cell = r[1]

        def glob(fr):
            v = cell[0]
            if v is UNSET:
                raise MLError("let rec value %s used before its definition" % name)
            return v
        return glob

    def is_prim(self, op, fs, scope):
        """Does the operator `op` still denote the built-in (so its arithmetic can be inlined)?"""
        if op not in self.prim_cells:
            return False
        s = fs
        while s is not None:
            if op in s.names:
                return False
            s = s.parent
        return scope.find_val(op) is self.prim_cells[op]

    def compile(self, e, fs, scope, name=None):
        kind = e[0]
        method = getattr(self, "c_" + kind, None)
        if method is None:
            raise MLError("unsupported expression %s" % kind)
        if kind in ("fun", "function"):
            return method(e, fs, scope, name)
        return method(e, fs, scope)

    def c_const(self, e, fs, scope):
        v = e[1]
        return lambda fr: v

    def c_var(self, e, fs, scope):
        return self.compile_var(e[1], fs, scope)

    def c_qvar(self, e, fs, scope):
        try:
            m = self.resolve_mod(scope, e[1])
        except MLError as ex:
            # a library this interpreter does not provide (Random, Gsl, ...): the function that
            # mentions it still compiles, and fails only if that line is executed
            msg = "%s (needed for %s.%s)" % (ex, ".".join(e[1]), e[2])

            def missing(fr):
                raise MLError(msg)
            return missing
        cell = m.vals.get(e[2])
        if cell is None:
            raise MLError("unbound value %s.%s" % (".".join(e[1]), e[2]))
        return lambda fr: cell[0]

    def c_open_in(self, e, fs, scope):
        inner = ModScope(scope, Module("<local open>"))
        inner.open(self.resolve_mod(scope, e[1]))
        return self.compile(e[2], fs, inner)

    def c_ctor(self, e, fs, scope):
        name = e[1]
        if e[2] is None:
            v = Ctor(name)
            return lambda fr: v
        a = self.compile(e[2], fs, scope)
        return lambda fr: Ctor(name, a(fr))

    def c_tuple(self, e, fs, scope):
        cs = [self.compile(x, fs, scope) for x in e[1]]
        return lambda fr: tuple([c(fr) for c in cs])

    def c_array(self, e, fs, scope):
        cs = [self.compile(x, fs, scope) for x in e[1]]
        return lambda fr: [c(fr) for c in cs]

    def c_list(self, e, fs, scope):
        cs = [self.compile(x, fs, scope) for x in e[1]]
        return lambda fr: from_list([c(fr) for c in cs])

    def c_cons(self, e, fs, scope):
        a = self.compile(e[1], fs, scope)
        b = self.compile(e[2], fs, scope)
        return lambda fr: Cons(a(fr), b(fr))

    def c_and(self, e, fs, scope):
        a = self.compile(e[1], fs, scope)
        b = self.compile(e[2], fs, scope)
        return lambda fr: a(fr) and b(fr)

    def c_or(self, e, fs, scope):
        a = self.compile(e[1], fs, scope)
        b = self.compile(e[2], fs, scope)
        return lambda fr: a(fr) or b(fr)

    def c_infix(self, e, fs, scope):
        op = e[1]
        a = self.compile(e[2], fs, scope)
        b = self.compile(e[3], fs, scope)
        if self.is_prim(op, fs, scope):
            if op in ("+.", "+"):
                return lambda fr: a(fr) + b(fr)
            if op in ("-.", "-"):
                return lambda fr: a(fr) - b(fr)
            if op in ("*.", "*"):
                return lambda fr: a(fr) * b(fr)
            if op == "/.":
                def div(fr):
                    x, y = a(fr), b(fr)
                    try:
                        return x / y
                    except ZeroDivisionError:
                        return fdiv(x, y)
                return div
            if op == "=":
                def eq(fr):
                    x, y = a(fr), b(fr)
                    return x == y if type(x) is float else equal(x, y)
                return eq
            if op == "<":
                def lt_(fr):
                    x, y = a(fr), b(fr)
                    return x < y if type(x) is float else lt(x, y)
                return lt_
            if op == ">":
                def gt_(fr):
                    x, y = a(fr), b(fr)
                    return x > y if type(x) is float else gt(x, y)
                return gt_
        f = self.compile_var(op, fs, scope)
        return lambda fr: apply(f(fr), [a(fr), b(fr)])

    def c_app(self, e, fs, scope):
        fe, args = e[1], e[2]
        # prefix primitives
        if fe[0] == "var" and len(args) == 1 and self.is_prim(fe[1], fs, scope):
            op = fe[1]
            a = self.compile(args[0], fs, scope)
            if op == "!":
                return lambda fr: a(fr).f["contents"]
            if op in ("~-.", "~-"):
                return lambda fr: -a(fr)
        f = self.compile(fe, fs, scope)
        cargs = []
        for x in args:
            if x[0] == "labarg":
                c = self.compile(x[2], fs, scope)
                cargs.append((lambda c, n, o: (lambda fr: Lab(n, c(fr), o)))(c, x[1], x[3]))
            else:
                cargs.append(self.compile(x, fs, scope))
        n = len(cargs)
        if n == 1:
            a0 = cargs[0]
            return lambda fr: apply(f(fr), [a0(fr)])
        if n == 2:
            a0, a1 = cargs
            return lambda fr: apply(f(fr), [a0(fr), a1(fr)])
        if n == 3:
            a0, a1, a2 = cargs
            return lambda fr: apply(f(fr), [a0(fr), a1(fr), a2(fr)])
        return lambda fr: apply(f(fr), [c(fr) for c in cargs])

    def c_if(self, e, fs, scope):
        c = self.compile(e[1], fs, scope)
        a = self.compile(e[2], fs, scope)
        if e[3] is None:
            def if1(fr):
                if c(fr):
                    a(fr)
                return UNIT
            return if1
        b = self.compile(e[3], fs, scope)
        return lambda fr: a(fr) if c(fr) else b(fr)

    def c_seq(self, e, fs, scope):
        # flatten a; b; c; ... into one loop
        parts = []
        while e[0] == "seq":
            parts.append(self.compile(e[1], fs, scope))
            e = e[2]
        last = self.compile(e, fs, scope)
        if len(parts) == 1:
            p0 = parts[0]

            def seq2(fr):
                p0(fr)
                return last(fr)
            return seq2

        def seq(fr):
            for p in parts:
                p(fr)
            return last(fr)
        return seq

    def c_let(self, e, fs, scope):
        _, rec, binds, body = e
        if rec:
            slots = []
            for pat, _ in binds:
                if pat[0] != "pvar":
                    raise MLError("let rec needs plain names")
                slots.append(fs.new_slot(pat[1]))
            cs = [self.compile(x, fs, scope, name=pat[1]) for pat, x in binds]
            cbody = self.compile(body, fs, scope)
            pairs = list(zip(slots, cs))

            def letrec(fr):
                for i, c in pairs:
                    fr[i] = c(fr)
                return cbody(fr)
            return letrec
        # let p1 = e1 and p2 = e2 in body: the e's do not see the p's
        saved = dict(fs.names)
        cs = [self.compile(x, fs, scope, name=pat[1] if pat[0] == "pvar" else None) for pat, x in binds]
        pms = [self.compile_pat(pat, fs) for pat, _ in binds]
        cbody = self.compile(body, fs, scope)
        fs.names = dict(saved)  # names the construct introduced go out of scope (slots stay theirs)
        if len(binds) == 1 and type(pms[0]) is int:
            i, c = pms[0], cs[0]

            def let1(fr):
                fr[i] = c(fr)
                return cbody(fr)
            return let1
        pairs = list(zip(pms, cs))

        def letn(fr):
            vals = [c(fr) for _, c in pairs]
            for (pm, _), v in zip(pairs, vals):
                if type(pm) is int:
                    fr[pm] = v
                elif not pm(fr, v):
                    raise match_failure("let")
            return cbody(fr)
        return letn

    def make_code(self, params, body, fs, scope, name):
        """params: parser param tuples.  Returns a Code whose body runs in a fresh FnScope."""
        inner = FnScope(fs)
        code = Code()
        code.name = name or "<fun>"
        code.simple = all(p[0] == "pos" for p in params)
        code.npos = len(params)
        binders, cparams = [], []
        for p in params:
            if p[0] == "pos":
                binders.append(self.compile_pat(p[1], inner))
                cparams.append(("pos", None))
            elif p[0] == "lab":
                binders.append(self.compile_pat(p[2], inner))
                cparams.append(("lab", p[1]))
            else:
                default = self.compile(p[3], inner, scope) if p[3] is not None else None
                binders.append(self.compile_pat(p[2], inner))
                cparams.append(("opt", p[1], None, default))
        code.binders = binders
        code.params = cparams
        code.body = self.compile(body, inner, scope)
        code.nslots = inner.nslots
        return code

    def c_fun(self, e, fs, scope, name=None):
        code = self.make_code(e[1], e[2], fs, scope, name)
        # `inner.nslots` is final only after the body is compiled, which make_code has done
        return lambda fr: Closure(code, fr)

    def c_function(self, e, fs, scope, name=None):
        arg = ("pvar", "%function-arg")
        body = ("match", ("var", "%function-arg"), e[1])
        code = self.make_code([("pos", arg)], body, fs, scope, name)
        return lambda fr: Closure(code, fr)

    def compile_arms(self, arms, fs, scope):
        out = []
        for pat, guard, body in arms:
            saved = dict(fs.names)
            pm = self.compile_pat_fn(pat, fs)
            g = self.compile(guard, fs, scope) if guard is not None else None
            b = self.compile(body, fs, scope)
            fs.names = dict(saved)  # names the construct introduced go out of scope (slots stay theirs)
            out.append((pm, g, b))
        return out

    def c_match(self, e, fs, scope):
        scrut = self.compile(e[1], fs, scope)
        arms = self.compile_arms(e[2], fs, scope)

        def match(fr):
            v = scrut(fr)
            for pm, g, b in arms:
                if pm(fr, v) and (g is None or g(fr)):
                    return b(fr)
            raise match_failure("match")
        return match

    def c_try(self, e, fs, scope):
        body = self.compile(e[1], fs, scope)
        arms = self.compile_arms(e[2], fs, scope)

        def try_(fr):
            try:
                return body(fr)
            except MLExn as ex:
                v = ex.value
                for pm, g, b in arms:
                    if pm(fr, v) and (g is None or g(fr)):
                        return b(fr)
                raise
        return try_

    def c_for(self, e, fs, scope):
        _, var, a, up, b, body = e
        ca = self.compile(a, fs, scope)
        cb = self.compile(b, fs, scope)
        saved = dict(fs.names)
        i = fs.new_slot(var)
        cbody = self.compile(body, fs, scope)
        fs.names = dict(saved)  # names the construct introduced go out of scope (slots stay theirs)
        if up:
            def for_up(fr):
                lo, hi = ca(fr), cb(fr)
                for k in range(lo, hi + 1):
                    fr[i] = k
                    cbody(fr)
                return UNIT
            return for_up

        def for_down(fr):
            hi, lo = ca(fr), cb(fr)
            for k in range(hi, lo - 1, -1):
                fr[i] = k
                cbody(fr)
            return UNIT
        return for_down

    def c_while(self, e, fs, scope):
        c = self.compile(e[1], fs, scope)
        body = self.compile(e[2], fs, scope)

        def while_(fr):
            while c(fr):
                body(fr)
            return UNIT
        return while_

    def c_field(self, e, fs, scope):
        a = self.compile(e[1], fs, scope)
        name = e[2]
        return lambda fr: a(fr).f[name]

    def c_setfield(self, e, fs, scope):
        a = self.compile(e[1], fs, scope)
        name = e[2]
        v = self.compile(e[3], fs, scope)

        def setfield(fr):
            a(fr).f[name] = v(fr)
            return UNIT
        return setfield

    def c_index(self, e, fs, scope):
        a = self.compile(e[1], fs, scope)
        i = self.compile(e[2], fs, scope)

        def index(fr):
            arr, k = a(fr), i(fr)
            if k < 0:
                raise MLExn(Ctor("Invalid_argument", "index out of bounds"))
            try:
                return arr[k]
            except IndexError:
                raise MLExn(Ctor("Invalid_argument", "index out of bounds"))
        return index

    def c_strindex(self, e, fs, scope):
        return self.c_index(e, fs, scope)

    def c_setindex(self, e, fs, scope):
        a = self.compile(e[1], fs, scope)
        i = self.compile(e[2], fs, scope)
        v = self.compile(e[3], fs, scope)

        def setindex(fr):
            arr, k, x = a(fr), i(fr), v(fr)
            if k < 0 or k >= len(arr):
                raise MLExn(Ctor("Invalid_argument", "index out of bounds"))
            arr[k] = x
            return UNIT
        return setindex

    def c_record(self, e, fs, scope):
        fields = [(f, self.compile(x, fs, scope)) for f, x in e[1]]
        if e[2] is not None:
            base = self.compile(e[2], fs, scope)

            def with_(fr):
                d = dict(base(fr).f)
                for f, c in fields:
                    d[f] = c(fr)
                return Record(d)
            return with_
        # keep the DECLARED field order (it is what Marshal and polymorphic compare see)
        order = self.record_types.get(fields[0][0])
        if order is not None and set(order) == {f for f, _ in fields}:
            rank = {f: k for k, f in enumerate(order)}
            evalorder = fields
            names = sorted((f for f, _ in fields), key=rank.__getitem__)

            def record_ordered(fr):
                d = {f: c(fr) for f, c in evalorder}
                return Record({f: d[f] for f in names})
            return record_ordered
        return lambda fr: Record({f: c(fr) for f, c in fields})

    def c_lazy(self, e, fs, scope):
        c = self.compile(e[1], fs, scope)
        return lambda fr: LazyVal(lambda: c(fr))

    def c_assert(self, e, fs, scope):
        c = self.compile(e[1], fs, scope)
        line = e[2]

        def assert_(fr):
            if not c(fr):
                raise MLExn(Ctor("Assert_failure", ("<file>", line, 0)))
            return UNIT
        return assert_


