"""CPU: the OCaml binding (ocaml/nbk.ml + ocaml/nbk_stubs.c) cannot be compiled in this image (no
OCaml toolchain), so its consistency with include/nbk.h is checked textually:
  * every integrator-facing symbol of include/nbk.h is called from a stub, with as many arguments
    as the header declares;
  * every `external` of nbk.ml names stubs that exist, whose C parameter count equals the arity of
    the OCaml type, with a bytecode twin whenever the arity exceeds five (OCaml's rule);
  * every CAMLprim stub is reachable from an `external`.
Device-pointer entry points (one rank of a torchrun job) and diagnostics are exempt: an OCaml
host is one process and uses nbk_create_multi."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EXEMPT = re.compile(r"^nbk_(dev_|peer_)|^nbk_(tune|sym_items|fp64_peak|last_kernel_ms|launch_count|sm_count|version|"
                    r"adv_range_profile|destroy|last_error)$")


def _strip_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _call_args(text, start):
    """text[start] is '(' : returns the argument list of that call."""
    depth, i = 0, start
    while True:
        if text[i] == "(":
            depth += 1
        elif text[i] == ")":
            depth -= 1
            if depth == 0:
                return _split_args(text[start + 1:i])
        i += 1


def header_functions():
    h = _strip_comments(open(os.path.join(ROOT, "include", "nbk.h")).read())
    fns = {}
    for m in re.finditer(r"\b(nbk_\w+)\s*\(", h):
        name = m.group(1)
        args = _call_args(h, m.end() - 1)
        fns[name] = 0 if args == ["void"] else len(args)
    return fns


def stub_text():
    return _strip_comments(open(os.path.join(ROOT, "ocaml", "nbk_stubs.c")).read())


def test_every_integrator_facing_symbol_is_bound_with_matching_arity():
    fns, c = header_functions(), stub_text()
    assert len(fns) >= 80
    missing, wrong = [], []
    for name, nargs in sorted(fns.items()):
        if EXEMPT.search(name):
            continue
        calls = [m for m in re.finditer(r"\b%s\s*\(" % re.escape(name), c)]
        if not calls:
            missing.append(name)
            continue
        for m in calls:
            got = _call_args(c, m.end() - 1)
            got = 0 if got == [] else len(got)
            if got != nargs:
                wrong.append((name, nargs, got))
    assert not missing, f"no OCaml stub calls: {missing}"
    assert not wrong, f"argument count differs from include/nbk.h: {wrong}"


def _ml_externals():
    ml = open(os.path.join(ROOT, "ocaml", "nbk.ml")).read()
    ml = re.sub(r"\(\*.*?\*\)", " ", ml, flags=re.S)
    out = []
    for m in re.finditer(r"external\s+(\w+)\s*:\s*(.*?)=\s*((?:\"\w+\"\s*)+)", ml, flags=re.S):
        name, typ, stubs = m.group(1), m.group(2), re.findall(r"\"(\w+)\"", m.group(3))
        # arity = top-level arrows (tuples and parenthesised types count as one argument)
        depth, arrows, i = 0, 0, 0
        while i < len(typ):
            if typ[i] == "(":
                depth += 1
            elif typ[i] == ")":
                depth -= 1
            elif typ.startswith("->", i) and depth == 0:
                arrows += 1
                i += 1
            i += 1
        out.append((name, arrows, stubs))
    return out


def _c_stubs():
    c = stub_text()
    out = {}
    for m in re.finditer(r"CAMLprim\s+value\s+(\w+)\s*\(", c):
        args = _call_args(c, m.end() - 1)
        out[m.group(1)] = args
    return out


def test_externals_match_their_stubs():
    ext, stubs = _ml_externals(), _c_stubs()
    assert len(ext) >= 50
    used = set()
    for name, arity, names in ext:
        assert arity >= 1, name
        if arity > 5:
            assert len(names) == 2, f"{name}: {arity} arguments need a bytecode and a native stub"
            bc, nat = names
            assert bc in stubs and re.match(r"value\s*\*", stubs[bc][0]) and len(stubs[bc]) == 2, (name, bc)
        else:
            assert len(names) == 1, name
            nat = names[0]
        assert nat in stubs, f"{name}: stub {nat} not defined in nbk_stubs.c"
        assert len(stubs[nat]) == arity, f"{name}: OCaml arity {arity}, C stub takes {len(stubs[nat])}"
        assert all(re.match(r"value\s+\w+$", a) for a in stubs[nat]), (name, stubs[nat])
        used.update(names)
    orphans = sorted(set(stubs) - used)
    assert not orphans, f"stubs without an external: {orphans}"


def test_glue_modules_only_use_declared_externals():
    """ocaml/*_gpu.ml call Nbk.<name>: every such name is an external or a helper of nbk.ml."""
    ml = open(os.path.join(ROOT, "ocaml", "nbk.ml")).read()
    declared = set(re.findall(r"external\s+(\w+)", ml)) | set(re.findall(r"^let\s+(\w+)", ml, flags=re.M))
    declared |= set(re.findall(r"^type\s+(\w+)", ml, flags=re.M))
    for fn in sorted(os.listdir(os.path.join(ROOT, "ocaml"))):
        if not fn.endswith("_gpu.ml"):
            continue
        src = open(os.path.join(ROOT, "ocaml", fn)).read()
        src = re.sub(r"\(\*.*?\*\)", " ", src, flags=re.S)
        for name in set(re.findall(r"\bNbk\.(\w+)", src)):
            assert name in declared, f"{fn} uses Nbk.{name}, which nbk.ml does not declare"
