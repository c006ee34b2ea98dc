"""CPU: pinning the oracle against the reference's own text.

The reference is OCaml and this image has no OCaml toolchain.  oracle/ref_fixtures/ holds small
OCaml programs (gen.ml, gen_wide.ml, gen_ics.ml, gen_dfloat.ml) that link the UNMODIFIED
reference modules and write what Base / Hermite / Leapfrog / OA / Advancer.A / Energy / Analysis /
Integrator / Ics / DFloat_hermite + DFloat_pushforward return for committed seeded inputs, as IEEE
bit patterns.  Two ways of running them:

  * oracle/miniml - an interpreter for the OCaml subset those sources use - executes them in this
    container straight from /root/reference (run_miniml.py); its outputs are committed under
    ref_miniml/ and the oracle must reproduce EVERY value BIT FOR BIT
    (test_oracle_reproduces_reference_outputs[ref_miniml]); with the reference tree present the
    interpreter is re-run and must reproduce the committed files
    (test_committed_ref_miniml_is_what_the_interpreter_produces).
  * `make -C oracle/ref_fixtures REF=...` on a machine with ocamlfind writes ref/*.hex from an
    ocamlopt build; [ref] compares against that once it exists and is skipped until then.

The other tests keep the kit honest: inputs and expected/ are exactly what the oracle produces
today, the generators emit the same keys, and the seeded inputs are the golden fixtures' own
bodies."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KIT = os.path.join(ROOT, "oracle", "ref_fixtures")
sys.path.insert(0, KIT)


def _kit():
    import make_inputs
    return make_inputs


def _parse(text):
    out = {}
    for ln in text.splitlines():
        key, i, j, k, bits = ln.split()
        out[(key, int(i), int(j), int(k))] = bits
    return out


def test_inputs_are_the_seeded_bodies_of_the_golden_fixtures(oracle):
    mi = _kit()
    m, q, p = mi.read_bodies(os.path.join(KIT, "inputs", "plummer_n64_seed20240.txt"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "sweeps_n64_eps0.npz"))
    assert np.array_equal(m, g["m"]) and np.array_equal(q, g["q"]) and np.array_equal(p, g["p"])
    m2, q2, p2 = oracle.make_plummer(64, 20240)
    assert np.array_equal(m, m2) and np.array_equal(q, q2) and np.array_equal(p, p2)
    m, q, p = mi.read_bodies(os.path.join(KIT, "inputs", "plummer_n16_std_seed20241.txt"))
    t = np.load(os.path.join(ROOT, "tests", "golden", "trajectories_n16.npz"))
    assert np.array_equal(m, t["m"]) and np.array_equal(q, t["q"]) and np.array_equal(p, t["p"])


def test_expected_files_are_what_the_oracle_returns_today(oracle):
    """expected/*.hex is regenerated from the compiled oracle and must match the committed text
    byte for byte (so `diff -r expected ref` on an OCaml machine compares against THIS oracle)."""
    for name, text in _kit().expected().items():
        assert open(os.path.join(KIT, "expected", name)).read() == text, name


def test_expected_agrees_with_the_golden_trajectories():
    t = np.load(os.path.join(ROOT, "tests", "golden", "trajectories_n16.npz"))
    e = _parse(open(os.path.join(KIT, "expected", "traj.hex")).read())
    bits = _kit().bits
    for i in range(16):
        for k in range(3):
            assert e[("hermite_q", i, 0, k)] == bits(t["hermite_q"][i, k])
            assert e[("oa_p", i, 0, k)] == bits(t["omelyan_p"][i, k])
            assert e[("lfc_v", i, 0, k)] == bits(t["leapfrog_v"][i, k])
            assert e[("lfa_q", i, 0, k)] == bits(t["leapfrog_adapt_q"][i, k])
            assert e[("adv_q", i, 0, k)] == bits(t["advancer_state"][i, 2 + k])


def test_gen_ml_emits_exactly_the_expected_keys_and_calls_the_reference():
    src = open(os.path.join(KIT, "gen.ml")).read()
    code = re.sub(r"\(\*.*?\*\)", " ", src, flags=re.S)
    for fn in ("Base.v ", "Base.grad_v ", "Base.grad_v_dgrad_v ", "Leapfrog.kick ", "Hermite.start_bs ",
               "Hermite.advance ", "Leapfrog.advance_const_dt ", "Leapfrog.advance ",
               "Omelyan_advancer.OA.advance ", "Advancer.A.advance ", "E.total_potential_energy ",
               "E.total_kinetic_energy ", "Energy.Make (Advancer.A)"):
        assert fn in code, fn
    keys = set(re.findall(r'emit3? oc "(\w+)"', code))
    keys |= {f"{k}{e}" for k in re.findall(r'Printf\.sprintf "(\w+)%d"', code) for e in (0, 1)}
    want = set()
    for name in ("pairs.hex", "sums.hex", "traj.hex"):
        want |= {k[0] for k in _parse(open(os.path.join(KIT, "expected", name)).read())}
    assert keys == want, (sorted(keys - want), sorted(want - keys))


SHEETS = ("pairs.hex", "sums.hex", "traj.hex", "wide.hex", "ics.hex", "dfloat.hex", "unit.hex")


@pytest.mark.parametrize("where", ["ref_miniml", "ref"])
def test_oracle_reproduces_reference_outputs(where):
    ref_dir = os.path.join(KIT, where)
    names = [n for n in SHEETS if os.path.exists(os.path.join(ref_dir, n))]
    if where == "ref" and not names:
        pytest.skip("oracle/ref_fixtures/ref/*.hex absent: no ocamlopt build of the generators has "
                    "been run (the interpreter-run pin is the [ref_miniml] case)")
    assert where == "ref" or names == list(SHEETS), "ref_miniml/ must hold every sheet"
    exp = _kit().expected()
    total = 0
    for n in names:
        ref, mine = _parse(open(os.path.join(ref_dir, n)).read()), _parse(exp[n])
        assert set(ref) == set(mine), n
        bad = [(k, ref[k], mine[k]) for k in sorted(ref) if ref[k] != mine[k]]
        assert not bad, f"{n}: {len(bad)} values differ from the reference, first: {bad[:5]}"
        total += len(ref)
    assert where == "ref" or total > 8000  # 8448 values at the time of writing


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree")
def test_committed_ref_miniml_is_what_the_interpreter_produces():
    """Re-run every generator over /root/reference with oracle/miniml (about 30 s) and compare with
    the committed ref_miniml/ byte for byte: the committed pin is reproducible, not hand-made."""
    import run_miniml
    for gen, files in run_miniml.GENERATORS.items():
        out = run_miniml.run_generator(gen, "/root/reference")
        for f in files:
            if f.endswith(".hex"):
                assert out[f] == open(os.path.join(KIT, "ref_miniml", f), "rb").read(), (gen, f)


def test_wide_generators_call_the_reference_and_cover_the_expected_keys():
    for gen, sheet, fns in (
            ("gen_wide.ml", "wide.hex",
             ("Base.set_eps ", "Hermite.start_bs ", "Hermite.advance ", "Advancer.A.advance ",
              "Advancer.A.advance ~extpot", "Omelyan_advancer.OA.advance ", "Leapfrog.advance ",
              "An.binaries ", "An.binaries_threshold ", "An.tightest_binary ",
              "An.bodies_to_el_phase_space ", "In.constant_sf_integrate ",
              "Analysis.Make (Advancer.A)", "Integrator.Make (Advancer.A) (Advancer.A)")),
            ("gen_ics.ml", "ics.hex", ("Ic.adjust_frame ", "Ic.rescale_to_standard_units ",
                                       "Ics.Make (Advancer.A)")),
            ("gen_dfloat.ml", "dfloat.hex", ("J.jacobian_advance ", "J.omega_pushforward ",
                                             "DFloat_pushforward.Make (DFloat_hermite)")),
            ("gen_unit.ml", "unit.hex", ("Hermite.advance ", "Leapfrog.advance ", "EL.energy ",
                                         "Energy.Make (Leapfrog)"))):
        code = re.sub(r"\(\*.*?\*\)", " ", open(os.path.join(KIT, gen)).read(), flags=re.S)
        for fn in fns:
            assert fn in code, (gen, fn)
        keys = {k[0] for k in _parse(open(os.path.join(KIT, "expected", sheet)).read())}
        stems = set(re.findall(r'"([a-z0-9_]+)"', code))
        for k in keys:  # every expected key is a literal of the generator, or stem ^ "_suffix"
            assert k in stems or any(k.startswith(st + "_") for st in stems), (gen, k)
