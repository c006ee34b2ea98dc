"""miniml - a small interpreter for the OCaml subset farr/direct-nbody's hot-path modules are
written in (TEST INFRASTRUCTURE; see README.md in this directory).

    from oracle.miniml import Interp   # repository root on sys.path
    it = Interp(search_path=["/path/to/direct-nbody"])
    base = it.load_file("/path/to/direct-nbody/base.ml")
"""
import sys
import threading

from .interp import (NIL, UNIT, Builtin, Closure, Cons, Ctor, Interp, MLError, MLExn, Record,  # noqa: F401
                     apply, call, from_list, to_list)
from .parser import ParseError, parse_program  # noqa: F401


def run_deep(fn, *args, stack_mb=512, recursion=2_000_000):
    """Run fn(*args) on a thread with a big stack: OCaml's tail calls are ordinary Python calls
    here, so a loop written as recursion nests as deep as it iterates."""
    box = {}

    def target():
        old = sys.getrecursionlimit()
        sys.setrecursionlimit(recursion)
        try:
            box["value"] = fn(*args)
        except BaseException as ex:  # noqa: BLE001 - re-raised on the caller's thread
            box["error"] = ex
        finally:
            sys.setrecursionlimit(old)
    prev = threading.stack_size(stack_mb << 20)
    try:
        th = threading.Thread(target=target)
        th.start()
        th.join()
    finally:
        threading.stack_size(prev)
    if "error" in box:
        raise box["error"]
    return box.get("value")
