#!/usr/bin/env python
"""Runs the reference-side fixture generators (gen.ml, gen_wide.ml, gen_ics.ml, gen_dfloat.ml) over
the UNMODIFIED reference sources with oracle/miniml - the small OCaml interpreter this repository
carries because its build image has no OCaml toolchain - and stores what they write under
ref_miniml/.  Test infrastructure.

    python oracle/ref_fixtures/run_miniml.py [--ref /root/reference] [--out DIR] [--only gen_wide]

The generators are the same files a maintainer compiles with ocamlfind (Makefile); nothing of the
reference is copied: its .ml files are read where they lie and interpreted.  ref_miniml/*.hex is
committed so that tests/test_ref_fixtures.py can compare the oracle with it wherever the tests
run; with the reference tree present (this container) the test re-runs the interpreter and
checks the committed files are what it produces.
"""
import argparse
import os
import shutil
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.miniml import Interp, run_deep  # noqa: E402

GENERATORS = {
    # generator -> files it writes under ref/
    "gen": ("pairs.hex", "sums.hex", "traj.hex", "states_adv.bin"),
    "gen_wide": ("wide.hex",),
    "gen_ics": ("ics.hex",),
    "gen_dfloat": ("dfloat.hex",),
    "gen_unit": ("unit.hex",),
}


def run_generator(name, ref, workdir=None, quiet=True):
    """Interpret HERE/<name>.ml against the reference tree `ref`; returns {file: bytes}."""
    tmp = workdir or tempfile.mkdtemp(prefix="miniml_")
    cwd = os.getcwd()
    try:
        if not os.path.exists(os.path.join(tmp, "inputs")):
            os.symlink(os.path.join(HERE, "inputs"), os.path.join(tmp, "inputs"))
        os.chdir(tmp)
        sink = open(os.devnull, "w") if quiet else None
        it = Interp(search_path=[ref, HERE, os.path.join(HERE, "diff_restated")],
                    stdout=sink, stderr=sink)
        run_deep(it.load_file, os.path.join(HERE, name + ".ml"))
        out = {}
        for f in GENERATORS[name]:
            with open(os.path.join(tmp, "ref", f), "rb") as fh:
                out[f] = fh.read()
        return out
    finally:
        os.chdir(cwd)
        if workdir is None:
            shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "ref_miniml"))
    ap.add_argument("--only", action="append")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for name in a.only or GENERATORS:
        t0 = time.time()
        for f, data in run_generator(name, a.ref).items():
            with open(os.path.join(a.out, f), "wb") as fh:
                fh.write(data)
        print("%-11s %6.1f s  -> %s" % (name, time.time() - t0, ", ".join(GENERATORS[name])))


if __name__ == "__main__":
    main()
