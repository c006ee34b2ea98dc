"""CPU: pinning the oracle against the REAL reference.

The reference is OCaml and this image has no OCaml toolchain, so nothing here can run it.
oracle/ref_fixtures/ is the kit for the first machine that can: gen.ml links the unmodified
reference modules and writes what Base / Hermite / Leapfrog / OA / Advancer.A / Energy return for
the committed seeded inputs (bit patterns); `make -C oracle/ref_fixtures REF=...` produces
ref/*.hex.  When those files exist, test_oracle_reproduces_reference_outputs asserts the oracle
reproduces every value BIT FOR BIT; until then it skips with PARITY UNPINNED, and the other tests
keep the kit honest: inputs and expected/ are exactly what the oracle produces today, gen.ml
emits the same keys, and the seeded inputs are the golden fixtures' own bodies."""
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


def test_oracle_reproduces_reference_outputs():
    ref_dir = os.path.join(KIT, "ref")
    names = ("pairs.hex", "sums.hex", "traj.hex")
    if not all(os.path.exists(os.path.join(ref_dir, n)) for n in names):
        pytest.skip("PARITY UNPINNED: oracle/ref_fixtures/ref/*.hex absent - no OCaml toolchain in "
                    "this image has ever run oracle/ref_fixtures/gen.ml against the reference")
    exp = _kit().expected()
    for n in names:
        ref, mine = _parse(open(os.path.join(ref_dir, n)).read()), _parse(exp[n])
        assert set(ref) == set(mine), n
        bad = [(k, ref[k], mine[k]) for k in sorted(ref) if ref[k] != mine[k]]
        assert not bad, f"{n}: {len(bad)} values differ from the reference, first: {bad[:5]}"
