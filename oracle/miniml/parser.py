"""miniml.parser - lexer and parser for the OCaml subset that farr/direct-nbody's hot-path modules
(base.ml, body.ml, energy.ml, hermite.ml, leapfrog.ml, advancer.ml, omelyan_advancer.ml, ...) and
oracle/ref_fixtures/gen.ml are written in.

TEST INFRASTRUCTURE (like everything under oracle/): it exists so that the UNMODIFIED reference
sources can be executed in an image that has no OCaml toolchain, which pins the C restatement in
oracle/ against what the reference's own text computes.  It is not a general OCaml front end:
types are skipped, not checked; objects, GADTs, first-class modules, polymorphic variants and
local modules are not parsed.

The grammar follows the precedence table of the OCaml manual (7.7 "Expressions"):
    !.. ~.. ?..  >  . .(  >  application  >  - -. (prefix)  >  **  >  * / % mod  >  + -  >  ::
    >  @ ^  >  = < > | & $ !=  >  & &&  >  or ||  >  ,  >  <- :=  >  if  >  ;  >  let match fun
"""
import re

KEYWORDS = {
    "and", "as", "assert", "begin", "do", "done", "downto", "else", "end", "exception", "external",
    "false", "for", "fun", "function", "functor", "if", "in", "include", "lazy", "let", "match",
    "module", "mutable", "of", "open", "rec", "sig", "struct", "then", "to", "true", "try", "type",
    "val", "when", "while", "with", "mod", "land", "lor", "lxor", "lsl", "lsr", "asr", "or",
}
OPCHARS = set("!$%&*+-./:<=>?@^|~")


class ParseError(Exception):
    pass


class Tok:
    __slots__ = ("kind", "val", "line")

    def __init__(self, kind, val, line):
        self.kind, self.val, self.line = kind, val, line

    def __repr__(self):
        return "%s:%r@%d" % (self.kind, self.val, self.line)


_num = re.compile(r"0[xX][0-9a-fA-F_]+|0[oO][0-7_]+|0[bB][01_]+|"
                  r"[0-9][0-9_]*(\.[0-9_]*)?([eE][+-]?[0-9][0-9_]*)?")
_ident = re.compile(r"[A-Za-z_][A-Za-z0-9_']*")
_char = re.compile(r"'(\\[\\\"'ntbr ]|\\[0-9]{3}|\\x[0-9a-fA-F]{2}|[^\\'])'")
_ESC = {"n": "\n", "t": "\t", "b": "\b", "r": "\r", "\\": "\\", '"': '"', "'": "'", " ": " "}


def _unescape(body, line):
    out, k = [], 0
    while k < len(body):
        c = body[k]
        if c != "\\":
            out.append(c)
            k += 1
            continue
        d = body[k + 1]
        if d in _ESC:
            out.append(_ESC[d])
            k += 2
        elif d.isdigit():
            out.append(chr(int(body[k + 1:k + 4])))
            k += 4
        elif d == "x":
            out.append(chr(int(body[k + 2:k + 4], 16)))
            k += 4
        elif d == "\n":  # line continuation: skip the newline and leading blanks
            k += 2
            while k < len(body) and body[k] in " \t":
                k += 1
        else:
            raise ParseError("line %d: bad escape \\%s" % (line, d))
    return "".join(out)


def lex(src):
    toks, k, line, n = [], 0, 1, len(src)
    while k < n:
        c = src[k]
        if c == "\n":
            line += 1
            k += 1
        elif c in " \t\r\f":
            k += 1
        elif c == "(" and src.startswith("(*", k) and not src.startswith("(*)", k):
            depth, k = 1, k + 2
            while depth:
                if k >= n:
                    raise ParseError("unterminated comment")
                if src.startswith("(*", k):
                    depth, k = depth + 1, k + 2
                elif src.startswith("*)", k):
                    depth, k = depth - 1, k + 2
                elif src[k] == '"':  # strings inside comments are lexed (OCaml does too)
                    k += 1
                    while k < n and src[k] != '"':
                        k += 2 if src[k] == "\\" else 1
                    k += 1
                else:
                    if src[k] == "\n":
                        line += 1
                    k += 1
        elif c == '"':
            j = k + 1
            while src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            body = src[k + 1:j]
            toks.append(Tok("string", _unescape(body, line), line))
            line += body.count("\n")
            k = j + 1
        elif c == "'":
            m = _char.match(src, k)
            if m:
                toks.append(Tok("char", _unescape(m.group(1), line), line))
                k = m.end()
            else:
                toks.append(Tok("op", "'", line))
                k += 1
        elif c.isdigit():
            m = _num.match(src, k)
            text = m.group(0)
            k = m.end()
            suffix = ""
            if k < n and src[k] in "lLn":
                suffix = src[k]
                k += 1
            clean = text.replace("_", "")
            if text[:2].lower() in ("0x", "0o", "0b"):
                toks.append(Tok("int", int(clean, 0), line))
            elif "." in text or "e" in text.lower():
                toks.append(Tok("float", float(clean), line))
            else:
                toks.append(Tok("int", int(clean), line))
            toks[-1].kind = toks[-1].kind if not suffix else "int"
        elif c.isalpha() or c == "_":
            m = _ident.match(src, k)
            w = m.group(0)
            k = m.end()
            if w in KEYWORDS:
                toks.append(Tok("kw", w, line))
            elif w == "_":
                toks.append(Tok("kw", "_", line))
            elif w[0].isupper():
                toks.append(Tok("uid", w, line))
            else:
                toks.append(Tok("lid", w, line))
        elif c == "[" and src.startswith("[|", k):
            toks.append(Tok("op", "[|", line))
            k += 2
        elif c == "|" and src.startswith("|]", k):
            toks.append(Tok("op", "|]", line))
            k += 2
        elif c in "()[]{},;":
            if src.startswith(";;", k):
                toks.append(Tok("op", ";;", line))
                k += 2
            else:
                toks.append(Tok("op", c, line))
                k += 1
        elif c in OPCHARS:
            j = k
            while j < n and src[j] in OPCHARS and not src.startswith("|]", j):
                j += 1
            # `x:=!y`, `a.(i)<-!r`, `x<-.5`: an assignment arrow ends its own token
            if j - k > 2 and src[k:k + 2] in (":=", "<-") and src[k + 2] in "!~-":
                j = k + 2
            toks.append(Tok("op", src[k:j], line))
            k = j
        else:
            raise ParseError("line %d: unexpected character %r" % (line, c))
    toks.append(Tok("eof", None, line))
    return toks


# ---------------------------------------------------------------------------------------------
# binary operators: (precedence, right-assoc) by spelling, OCaml manual 7.7
def infix_prec(tok):
    if tok.kind == "kw":
        v = tok.val
        if v == "or":
            return (1, True)
        if v in ("mod", "land", "lor", "lxor"):
            return (7, False)
        if v in ("lsl", "lsr", "asr"):
            return (8, True)
        return None
    if tok.kind != "op":
        return None
    v = tok.val
    if v in ("||",):
        return (1, True)
    if v in ("&&", "&"):
        return (2, True)
    if v in ("<-", ":=", "->", "|", ";", ";;", ",", ".", "'", ":", "::", "!", "?", "~", "[|", "|]"):
        return (5, True) if v == "::" else None
    c = v[0]
    if v.startswith("**"):
        return (8, True)
    if c in "=<>|&$" or v == "!=":
        return (3, False)
    if c in "@^":
        return (4, True)
    if c in "+-":
        return (6, False)
    if c in "*/%":
        return (7, False)
    return None


class Parser:
    def __init__(self, src, name="<src>"):
        self.toks = lex(src)
        self.k = 0
        self.name = name

    # -- token helpers
    @property
    def tok(self):
        return self.toks[self.k]

    def peek(self, d=1):
        return self.toks[min(self.k + d, len(self.toks) - 1)]

    def err(self, msg):
        t = self.tok
        raise ParseError("%s:%d: %s (at %r)" % (self.name, t.line, msg, t.val))

    def at(self, val, kind=None):
        t = self.tok
        return t.val == val and t.kind in (("op", "kw") if kind is None else (kind,))

    def accept(self, val):
        if self.at(val):
            self.k += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            self.err("expected %r" % val)

    def next(self):
        t = self.tok
        self.k += 1
        return t

    # -- types are skipped ---------------------------------------------------------------------
    def skip_type(self, stops):
        """Skip a type expression: stops at a depth-0 token whose value is in `stops`."""
        depth = 0
        while True:
            t = self.tok
            if t.kind == "eof":
                return
            if t.kind in ("op", "kw"):
                if depth == 0 and t.val in stops:
                    return
                if t.val in ("(", "[", "{", "[|"):
                    depth += 1
                elif t.val in (")", "]", "}", "|]"):
                    if depth == 0:
                        return
                    depth -= 1
            self.k += 1

    TOPLEVEL_STARTS = ("let", "type", "module", "open", "include", "exception", "external", "val",
                       "end", ";;", "and")

    # -- structure items -------------------------------------------------------------------------
    def parse_structure(self, until_end=False):
        items = []
        while True:
            t = self.tok
            if t.kind == "eof":
                if until_end:
                    self.err("missing 'end'")
                return items
            if until_end and self.at("end"):
                return items
            if self.accept(";;"):
                continue
            items.append(self.parse_item())

    def parse_item(self):
        t = self.tok
        line = t.line
        if self.at("let"):
            self.k += 1
            rec = self.accept("rec")
            binds = self.parse_bindings()
            if self.accept("in"):
                body = self.parse_seq()
                return ("eval", ("let", rec, binds, body), line)
            return ("let", rec, binds, line)
        if self.at("type"):
            self.k += 1
            return ("type", self.parse_typedefs(), line)
        if self.at("exception"):
            self.k += 1
            name = self.next().val
            if self.accept("of"):
                self.skip_type(self.TOPLEVEL_STARTS)
            elif self.accept("="):
                self.parse_modpath()
            return ("exception", name, line)
        if self.at("open"):
            self.k += 1
            self.accept("!")
            return ("open", self.parse_modpath(), line)
        if self.at("include"):
            self.k += 1
            return ("include", self.parse_modexpr(), line)
        if self.at("external"):
            self.k += 1
            name = self.parse_value_name()
            self.expect(":")
            self.skip_type(("=",))
            self.expect("=")
            prims = []
            while self.tok.kind == "string":
                prims.append(self.next().val)
            return ("external", name, prims, line)
        if self.at("module"):
            self.k += 1
            if self.accept("type"):
                self.next()  # name
                if self.accept("="):
                    self.skip_modtype()
                return ("modtype", line)
            name = self.next().val
            params = []
            while self.at("("):
                self.k += 1
                pname = self.next().val
                self.expect(":")
                self.skip_modtype()
                self.expect(")")
                params.append(pname)
            if self.accept(":"):
                self.skip_modtype()
            self.expect("=")
            body = self.parse_modexpr()
            if params:
                body = ("functor", params, body)
            return ("module", name, body, line)
        # a bare top-level expression
        return ("eval", self.parse_seq(), line)

    def parse_modpath(self):
        path = [self.next().val]
        while self.at(".") and self.peek().kind == "uid":
            self.k += 1
            path.append(self.next().val)
        return path

    def parse_modexpr(self):
        if self.at("struct"):
            self.k += 1
            items = self.parse_structure(until_end=True)
            self.expect("end")
            return ("struct", items)
        if self.at("functor"):
            self.k += 1
            params = []
            while self.accept("("):
                params.append(self.next().val)
                self.expect(":")
                self.skip_modtype()
                self.expect(")")
            self.expect("->")
            return ("functor", params, self.parse_modexpr())
        if self.at("("):
            self.k += 1
            me = self.parse_modexpr()
            if self.accept(":"):
                self.skip_modtype()
            self.expect(")")
        else:
            if self.tok.kind != "uid":
                self.err("module expression expected")
            me = ("mpath", self.parse_modpath())
        while self.at("("):  # functor application
            self.k += 1
            arg = self.parse_modexpr()
            if self.accept(":"):
                self.skip_modtype()
            self.expect(")")
            me = ("mapply", me, arg)
        return me

    def skip_modtype(self):
        """Skip a module type: sig ... end | path | (mt) | functor(..) -> mt, then `with ...`."""
        if self.at("sig"):
            depth = 0
            while True:
                t = self.next()
                if t.kind == "eof":
                    self.err("unterminated sig")
                if t.kind == "kw" and t.val in ("sig", "struct", "begin"):
                    depth += 1
                elif t.kind == "kw" and t.val == "end":
                    depth -= 1
                    if depth == 0:
                        break
        elif self.at("("):
            self.k += 1
            self.skip_modtype()
            self.expect(")")
        elif self.at("functor"):
            self.k += 1
            while self.accept("("):
                self.next()
                self.expect(":")
                self.skip_modtype()
                self.expect(")")
            self.expect("->")
            self.skip_modtype()
        else:
            self.parse_modpath()
            if self.at(".") and self.peek().kind == "lid":  # path ending in a lowercase type name
                self.k += 2
        while self.at("with"):
            self.k += 1
            while True:
                if not (self.accept("type") or self.accept("module")):
                    self.err("with-constraint expected")
                self.skip_type(("=", ":="))
                self.next()
                self.skip_type(("and", ")", "=", "with") + self.TOPLEVEL_STARTS)
                if not self.accept("and"):
                    break

    def parse_typedefs(self):
        """type [params] name = {fields} | variants | alias  [and ...]; returns what the
        interpreter needs: [('record', name, [(field, mutable)]) | ('variant', name, [(ctor, has_arg)])
        | ('other', name)]"""
        defs = []
        while True:
            # optional type parameters: 'a  or ('a, 'b)
            if self.at("'"):
                self.k += 2
            elif self.at("(") and self.peek().val == "'":
                self.skip_type((")",))
                self.expect(")")
            name = self.next().val
            if not self.accept("="):
                defs.append(("other", name))
            elif self.at("{"):
                self.k += 1
                fields = []
                while not self.at("}"):
                    mut = self.accept("mutable")
                    fname = self.next().val
                    self.expect(":")
                    self.skip_type((";", "}"))
                    fields.append((fname, mut))
                    self.accept(";")
                self.expect("}")
                defs.append(("record", name, fields))
            elif self.at("|") or (self.tok.kind == "uid" and not (self.peek().val == ".")):
                self.accept("|")
                ctors = []
                while True:
                    cname = self.next().val
                    has_arg = False
                    if self.accept("of"):
                        has_arg = True
                        self.skip_type(("|", "and") + self.TOPLEVEL_STARTS)
                    ctors.append((cname, has_arg))
                    if not self.accept("|"):
                        break
                defs.append(("variant", name, ctors))
            else:
                self.skip_type(("and",) + self.TOPLEVEL_STARTS)
                defs.append(("other", name))
            if not self.accept("and"):
                return defs

    # -- let bindings ------------------------------------------------------------------------------
    def parse_value_name(self):
        if self.at("("):
            self.k += 1
            t = self.next()
            self.expect(")")
            return t.val
        return self.next().val

    def parse_bindings(self):
        binds = [self.parse_binding()]
        while self.accept("and"):
            binds.append(self.parse_binding())
        return binds

    def starts_param(self):
        t = self.tok
        if t.kind in ("lid", "uid", "int", "float", "string", "char"):
            return True
        return t.val in ("(", "[", "[|", "{", "_", "~", "?", "true", "false") and t.kind in ("op", "kw")

    def parse_binding(self):
        # `name params = e`  (function)   or   `pattern = e`
        t = self.tok
        is_fn_name = (t.kind == "lid" or (t.val == "(" and self.peek().kind in ("op", "kw")
                                          and self.peek(2).val == ")" and self.peek().val not in ("(", ")")))
        if is_fn_name:
            save = self.k
            name = self.parse_value_name()
            if self.at("=") or self.at(":"):
                if self.accept(":"):
                    self.skip_type(("=",))
                self.expect("=")
                return (("pvar", name), self.parse_seq())
            if self.starts_param():
                params = []
                while self.starts_param():
                    params.append(self.parse_param())
                if self.accept(":"):
                    self.skip_type(("=",))
                self.expect("=")
                return (("pvar", name), ("fun", params, self.parse_seq()))
            self.k = save
        pat = self.parse_pattern()
        if self.accept(":"):
            self.skip_type(("=",))
        self.expect("=")
        return (pat, self.parse_seq())

    def parse_param(self):
        if self.accept("~"):
            if self.accept("("):
                name = self.next().val
                if self.accept(":"):
                    self.skip_type((")",))
                self.expect(")")
                return ("lab", name, ("pvar", name))
            name = self.next().val
            if self.accept(":"):
                return ("lab", name, self.parse_simple_pattern())
            return ("lab", name, ("pvar", name))
        if self.accept("?"):
            if self.accept("("):
                name = self.next().val
                if self.accept(":"):
                    self.skip_type((")", "="))
                default = self.parse_seq() if self.accept("=") else None
                self.expect(")")
                return ("opt", name, ("pvar", name), default)
            name = self.next().val
            if self.accept(":"):
                if self.accept("("):
                    pat = self.parse_pattern()
                    if self.accept(":"):
                        self.skip_type((")", "="))
                    default = self.parse_seq() if self.accept("=") else None
                    self.expect(")")
                    return ("opt", name, pat, default)
                return ("opt", name, self.parse_simple_pattern(), None)
            return ("opt", name, ("pvar", name), None)
        return ("pos", self.parse_simple_pattern())

    # -- patterns ----------------------------------------------------------------------------------
    def parse_pattern(self):
        p = self.parse_pattern_or()
        while self.accept("as"):
            p = ("pas", p, self.next().val)
        return p

    def parse_pattern_or(self):
        p = self.parse_pattern_tuple()
        while self.at("|"):
            self.k += 1
            p = ("por", p, self.parse_pattern_tuple())
        return p

    def parse_pattern_tuple(self):
        p = self.parse_pattern_cons()
        if self.at(","):
            ps = [p]
            while self.accept(","):
                ps.append(self.parse_pattern_cons())
            return ("ptuple", ps)
        return p

    def parse_pattern_cons(self):
        p = self.parse_pattern_app()
        if self.accept("::"):
            return ("pcons", p, self.parse_pattern_cons())
        return p

    def starts_simple_pattern(self):
        t = self.tok
        if t.kind in ("lid", "uid", "int", "float", "string", "char"):
            return True
        return t.kind in ("op", "kw") and t.val in ("(", "[", "[|", "{", "_", "true", "false", "-", "-.")

    def parse_pattern_app(self):
        if self.tok.kind == "uid":
            path = self.parse_modpath()
            if self.at(".") and self.peek().val in ("(", "{", "["):  # M.(pattern): not needed
                self.err("module-qualified pattern")
            name = path[-1]
            if self.starts_simple_pattern():
                return ("pctor", name, self.parse_simple_pattern())
            return ("pctor", name, None)
        return self.parse_simple_pattern()

    def parse_simple_pattern(self):
        t = self.tok
        if t.kind == "lid":
            self.k += 1
            return ("pvar", t.val)
        if t.kind == "uid":
            path = self.parse_modpath()
            return ("pctor", path[-1], None)
        if t.kind in ("int", "float", "string", "char"):
            self.k += 1
            return ("pconst", t.val)
        if t.val in ("-", "-.") and t.kind == "op":
            self.k += 1
            return ("pconst", -self.next().val)
        if self.accept("_"):
            return ("pany",)
        if self.accept("true"):
            return ("pconst", True)
        if self.accept("false"):
            return ("pconst", False)
        if self.accept("("):
            if self.accept(")"):
                return ("punit",)
            # an operator name in a binding position: let ( +. ) = ...
            if self.tok.kind in ("op", "kw") and self.peek().val == ")" and infix_prec(self.tok):
                name = self.next().val
                self.expect(")")
                return ("pvar", name)
            p = self.parse_pattern()
            if self.accept(":"):
                self.skip_type((")",))
            self.expect(")")
            return p
        if self.accept("["):
            ps = []
            while not self.at("]"):
                ps.append(self.parse_pattern())
                if not self.accept(";"):
                    break
            self.expect("]")
            out = ("pnil",)
            for p in reversed(ps):
                out = ("pcons", p, out)
            return out
        if self.accept("[|"):
            ps = []
            while not self.at("|]"):
                ps.append(self.parse_pattern())
                if not self.accept(";"):
                    break
            self.expect("|]")
            return ("parray", ps)
        if self.accept("{"):
            fields = []
            while not self.at("}"):
                if self.accept("_"):
                    self.accept(";")
                    continue
                path = []
                while self.tok.kind == "uid":
                    path.append(self.next().val)
                    self.expect(".")
                fname = self.next().val
                if self.accept("="):
                    fields.append((fname, self.parse_pattern()))
                else:
                    fields.append((fname, ("pvar", fname)))
                if not self.accept(";"):
                    break
            self.expect("}")
            return ("precord", fields)
        self.err("pattern expected")

    # -- expressions --------------------------------------------------------------------------------
    def parse_seq(self):
        e = self.parse_expr()
        if self.at(";"):
            self.k += 1
            # a trailing `;` before a closer is allowed
            if self.at("end") or self.at("done") or self.at(")") or self.at("|") or self.at("}") \
                    or self.at(";;") or self.tok.kind == "eof" or self.at("]") or self.at("|]"):
                return e
            return ("seq", e, self.parse_seq())
        return e

    def parse_expr(self):
        """One expression without a top-level `;`."""
        if self.at("let") or self.at("match") or self.at("fun") or self.at("function") \
                or self.at("try") or self.at("if"):
            return self.parse_open_form()
        lhs = self.parse_tuple()
        if self.at("<-"):
            self.k += 1
            rhs = self.parse_expr()
            if lhs[0] == "field":
                return ("setfield", lhs[1], lhs[2], rhs)
            if lhs[0] == "index":
                return ("setindex", lhs[1], lhs[2], rhs)
            self.err("bad left-hand side of <-")
        if self.at(":="):
            self.k += 1
            return ("app", ("var", ":="), [lhs, self.parse_expr()])
        return lhs

    def parse_open_form(self):
        t = self.tok
        if self.accept("let"):
            if self.accept("open"):
                path = self.parse_modpath()
                self.expect("in")
                return ("open_in", path, self.parse_seq())
            rec = self.accept("rec")
            binds = self.parse_bindings()
            self.expect("in")
            return ("let", rec, binds, self.parse_seq())
        if self.accept("if"):
            c = self.parse_seq()
            self.expect("then")
            a = self.parse_expr()
            b = self.parse_expr() if self.accept("else") else None
            return ("if", c, a, b)
        if self.accept("match"):
            e = self.parse_seq()
            self.expect("with")
            return ("match", e, self.parse_arms())
        if self.accept("try"):
            e = self.parse_seq()
            self.expect("with")
            return ("try", e, self.parse_arms())
        if self.accept("function"):
            return ("function", self.parse_arms())
        if self.accept("fun"):
            params = []
            while not self.at("->"):
                params.append(self.parse_param())
            self.expect("->")
            return ("fun", params, self.parse_seq())
        self.err("expression expected (%s)" % t.val)

    def parse_arms(self):
        self.accept("|")
        arms = []
        while True:
            pat = self.parse_pattern()
            guard = self.parse_seq() if self.accept("when") else None
            self.expect("->")
            arms.append((pat, guard, self.parse_seq()))
            if not self.accept("|"):
                return arms

    def parse_tuple(self):
        e = self.parse_binop(1)
        if self.at(","):
            es = [e]
            while self.accept(","):
                es.append(self.parse_operand_or_open(1))
            return ("tuple", es)
        return e

    def parse_operand_or_open(self, prec):
        if self.at("let") or self.at("match") or self.at("fun") or self.at("function") \
                or self.at("try") or self.at("if"):
            return self.parse_open_form()
        return self.parse_binop(prec)

    def parse_binop(self, min_prec):
        lhs = self.parse_unary()
        while True:
            t = self.tok
            pa = infix_prec(t)
            if pa is None or pa[0] < min_prec:
                return lhs
            prec, right = pa
            self.k += 1
            if self.at("let") or self.at("match") or self.at("fun") or self.at("function") \
                    or self.at("try") or self.at("if"):
                rhs = self.parse_open_form()
            else:
                rhs = self.parse_binop(prec if right else prec + 1)
            op = t.val
            if op == "::":
                lhs = ("cons", lhs, rhs)
            elif op in ("&&", "&"):
                lhs = ("and", lhs, rhs)
            elif op in ("||", "or"):
                lhs = ("or", lhs, rhs)
            else:
                lhs = ("infix", op, lhs, rhs)

    def parse_unary(self):
        t = self.tok
        if t.kind == "op" and t.val in ("-", "-.", "+", "+."):
            self.k += 1
            if self.at("let") or self.at("match") or self.at("fun") or self.at("function") \
                    or self.at("try") or self.at("if"):
                e = self.parse_open_form()
            else:
                e = self.parse_binop(8)  # binds looser than ** only
            if t.val in ("+", "+."):
                return e
            if e[0] == "const" and isinstance(e[1], (int, float)) and not isinstance(e[1], bool):
                return ("const", -e[1])
            return ("app", ("var", "~-" if t.val == "-" else "~-."), [e])
        return self.parse_app()

    def starts_arg(self):
        t = self.tok
        if t.kind in ("lid", "uid", "int", "float", "string", "char"):
            return True
        if t.kind == "kw":
            return t.val in ("begin", "true", "false", "for", "while")
        if t.kind == "op":
            v = t.val
            if v in ("(", "[", "[|", "{", "~", "?"):
                return True
            return v[0] in "!" and v != "!=" or (v[0] in "~?" and len(v) > 1)
        return False

    def parse_app(self):
        if self.accept("assert"):
            return ("assert", self.parse_arg(), self.toks[self.k - 1].line)
        if self.accept("lazy"):
            return ("lazy", self.parse_arg())
        t = self.tok
        # constructor application: Uid [arg]   (but Uid.x is a module path)
        if t.kind == "uid":
            save = self.k
            path = self.parse_modpath()
            if not self.at("."):
                arg = self.parse_arg() if self.starts_arg() and not self.tok.val in ("~", "?") else None
                return ("ctor", path[-1], arg)
            self.k = save
        f = self.parse_arg()
        args = []
        while self.starts_arg():
            args.append(self.parse_arg(labels=True))
        return ("app", f, args) if args else f

    def parse_arg(self, labels=False):
        t = self.tok
        if labels and t.kind == "op" and t.val in ("~", "?") and self.peek().kind == "lid":
            self.k += 1
            name = self.next().val
            if self.at(":") :
                self.k += 1
                return ("labarg", name, self.parse_arg(), t.val == "?")
            return ("labarg", name, ("var", name), t.val == "?")
        if t.kind == "op" and t.val[0] in "!~?" and t.val not in ("!=", "~", "?"):
            # prefix symbols bind tighter than . and .( : `!r.(0)` is `(!r).(0)`
            return self.parse_postfix(self.parse_prefixed())
        return self.parse_postfix(self.parse_atom())

    def parse_prefixed(self):
        t = self.tok
        if t.kind == "op" and t.val[0] in "!~?" and t.val not in ("!=", "~", "?"):
            self.k += 1
            return ("app", ("var", t.val), [self.parse_prefixed()])
        return self.parse_atom()

    def parse_postfix(self, e):
        while self.at("."):
            nxt = self.peek()
            if nxt.val == "(" and nxt.kind == "op":
                self.k += 2
                i = self.parse_seq()
                self.expect(")")
                e = ("index", e, i)
            elif nxt.val == "[" and nxt.kind == "op":
                self.k += 2
                i = self.parse_seq()
                self.expect("]")
                e = ("strindex", e, i)
            elif nxt.val == "{" and nxt.kind == "op":  # Bigarray.Array1: a.{i}
                self.k += 2
                i = self.parse_seq()
                self.expect("}")
                e = ("index", e, i)
            elif nxt.kind == "lid":
                self.k += 2
                e = ("field", e, nxt.val)
            elif nxt.kind == "uid":  # b.Module.field
                self.k += 1
                self.parse_modpath()
                self.expect(".")
                e = ("field", e, self.next().val)
            else:
                self.err("bad postfix")
        return e

    def parse_atom(self):
        t = self.tok
        if t.kind in ("int", "float", "string", "char"):
            self.k += 1
            return ("const", t.val)
        if t.kind == "lid":
            self.k += 1
            return ("var", t.val)
        if t.kind == "uid":
            path = self.parse_modpath()
            if self.accept("."):
                if self.at("("):
                    # Module.(op)  or  Module.(expr)
                    if self.peek().kind in ("op", "kw") and self.peek(2).val == ")" \
                            and self.peek().val not in ("(", ")"):
                        self.k += 1
                        op = self.next().val
                        self.expect(")")
                        return ("qvar", path, op)
                    self.k += 1
                    e = self.parse_seq()
                    self.expect(")")
                    return ("open_in", path, e)
                name = self.next()
                if name.kind != "lid":
                    self.err("value name expected after module path")
                return ("qvar", path, name.val)
            return ("ctor", path[-1], None)
        if self.accept("true"):
            return ("const", True)
        if self.accept("false"):
            return ("const", False)
        if self.accept("("):
            if self.accept(")"):
                return ("const", ())
            # operator section: ( +. )  ( * )  (mod)
            if self.tok.kind in ("op", "kw") and self.peek().val == ")" and self.peek().kind == "op" \
                    and (infix_prec(self.tok) or self.tok.val in ("!", "~-", "~-.", ":=")):
                op = self.next().val
                self.expect(")")
                return ("var", op)
            e = self.parse_seq()
            if self.accept(":"):
                self.skip_type((")",))
            self.expect(")")
            return e
        if self.accept("begin"):
            if self.accept("end"):
                return ("const", ())
            e = self.parse_seq()
            self.expect("end")
            return e
        if self.accept("["):
            es = []
            while not self.at("]"):
                es.append(self.parse_expr())
                if not self.accept(";"):
                    break
            self.expect("]")
            return ("list", es)
        if self.accept("[|"):
            es = []
            while not self.at("|]"):
                es.append(self.parse_expr())
                if not self.accept(";"):
                    break
            self.expect("|]")
            return ("array", es)
        if self.accept("{"):
            base = None
            # { e with f = v; ... }: try a leading expression followed by `with`
            save = self.k
            if not (self.tok.kind == "lid" and self.peek().val in ("=", ";", "}")) \
                    and not (self.tok.kind == "uid" and self.peek().val == "."
                             and self._qualified_field_then_eq()):
                base = self.parse_arg()
                self.expect("with")
            else:
                self.k = save
            fields = []
            while not self.at("}"):
                while self.tok.kind == "uid":
                    self.k += 1
                    self.expect(".")
                fname = self.next().val
                if self.accept("="):
                    fields.append((fname, self.parse_expr()))
                else:
                    fields.append((fname, ("var", fname)))
                if not self.accept(";"):
                    break
            self.expect("}")
            return ("record", fields, base)
        if self.accept("for"):
            var = self.next().val
            self.expect("=")
            a = self.parse_seq()
            if self.accept("to"):
                up = True
            else:
                self.expect("downto")
                up = False
            b = self.parse_seq()
            self.expect("do")
            body = ("const", ()) if self.at("done") else self.parse_seq()
            self.expect("done")
            return ("for", var, a, up, b, body)
        if self.accept("while"):
            c = self.parse_seq()
            self.expect("do")
            body = ("const", ()) if self.at("done") else self.parse_seq()
            self.expect("done")
            return ("while", c, body)
        self.err("expression expected")

    def _qualified_field_then_eq(self):
        """At `Uid . ... lid`: is this a module-qualified FIELD LABEL followed by `=` (a record
        construction) rather than an expression such as `Module.value with ...`?"""
        j = self.k
        while self.toks[j].kind == "uid" and self.toks[j + 1].val == ".":
            j += 2
        return self.toks[j].kind == "lid" and self.toks[j + 1].val == "="


def parse_program(src, name="<src>"):
    p = Parser(src, name)
    return p.parse_structure()
