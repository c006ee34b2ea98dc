"""CPU: the `--- !Particle` text snapshots (A.print / A.read advancer.ml:88-96,118-124;
Leapfrog.print / read leapfrog.ml:44-62) written and read by the host mirror (libnbh), against
the oracle's restatement and a committed golden file.  Byte-exact: the format is text."""
import os

import numpy as np
import pytest

import nbh

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bodies(oracle, n=37, seed=7):
    m, q, p = oracle.make_plummer(n, seed)
    m = m * np.linspace(0.5, 3.0, n)          # unequal masses
    ids = np.arange(100, 100 + n, dtype=np.int32)
    t = np.full(n, 0.125) + 1e-3 * np.arange(n)
    q[0] = [1e-7, -2.5e300, 0.0]              # %g corner cases: exponents, zero
    p[1] = [123456789.0 * m[1], -1e-5 * m[1], 1.5 * m[1]]
    return ids, t, m, q, p


def test_advancer_print_is_byte_identical_to_the_restatement(tmp_path, oracle):
    ids, t, m, q, p = _bodies(oracle)
    path = tmp_path / "a.yaml"
    nbh.advancer_print(str(path), ids, t, m, q, p)
    assert path.read_text() == oracle.advancer_print_text(ids, t, m, q, p)
    # append mode: the reference prints body after body to one channel
    nbh.advancer_print(str(path), ids[:3], t[:3], m[:3], q[:3], p[:3], append=True)
    assert path.read_text() == (oracle.advancer_print_text(ids, t, m, q, p) +
                                oracle.advancer_print_text(ids[:3], t[:3], m[:3], q[:3], p[:3]))


def test_advancer_read_round_trip_and_restatement(tmp_path, oracle):
    ids, t, m, q, p = _bodies(oracle)
    path = tmp_path / "a.yaml"
    nbh.advancer_print(str(path), ids, t, m, q, p)
    got = nbh.advancer_read(str(path))
    ref = oracle.advancer_read_text(path.read_text())
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)            # same parse, same p = v * m product
    assert np.array_equal(got[0], ids)
    assert np.allclose(got[2], m, rtol=1e-5) and np.allclose(got[3], q, rtol=1e-5)
    assert np.allclose(got[4], p, rtol=2e-5)   # %g keeps 6 significant digits
    # what was read prints to the same text again (fixed point of print . read)
    path2 = tmp_path / "b.yaml"
    nbh.advancer_print(str(path2), *got)
    again = nbh.advancer_read(str(path2))
    assert np.array_equal(again[3], got[3]) and np.array_equal(again[2], got[2])


def test_leapfrog_print_read(tmp_path, oracle):
    ids, t, m, q, p = _bodies(oracle, 21, 9)
    v = p / m[:, None]
    dtm = np.abs(q[:, 0]) * 1e-2
    dtm[3] = np.inf                            # a body that has not been kicked yet
    path = tmp_path / "l.yaml"
    nbh.leapfrog_print(str(path), ids, t, dtm, m, q, v)
    text = path.read_text()
    assert text == oracle.leapfrog_print_text(ids, t, dtm, m, q, v)
    assert "t_max: inf\n" in text
    got = nbh.leapfrog_read(str(path))
    ref = oracle.leapfrog_read_text(text)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    assert np.isinf(got[2][3]) and np.allclose(got[5], v, rtol=1e-5)


def test_golden_snapshot_file(oracle):
    """tests/golden/particles_advancer.yaml was written by the oracle restatement
    (tests/golden/make_golden.py); the host mirror must reproduce it byte for byte from the
    values it reads out of it."""
    path = os.path.join(GOLDEN, "particles_advancer.yaml")
    text = open(path).read()
    got = nbh.advancer_read(path)
    assert len(got[0]) == 5 and list(got[0]) == [0, 1, 2, 3, 4]
    assert oracle.advancer_print_text(*got) == text


def test_reader_errors_like_scan_failure(tmp_path):
    bad = tmp_path / "bad.yaml"
    bad.write_text("--- !Particle\nid: 1\nt: 0\nr:\n  - 1\n  - 2\nv:\n")
    with pytest.raises(nbh.NbhError):
        nbh.advancer_read(str(bad))
    with pytest.raises(nbh.NbhError):
        nbh.advancer_read(str(tmp_path / "missing.yaml"))
    ok = tmp_path / "ok.yaml"   # free-form white space, as Scanf accepts it
    ok.write_text("---   !Particle id: 7 t: 1e-3 r: - 1 - 2 - 3 v: - -4 -5 - 6\n m: 2")
    ids, t, m, q, p = nbh.advancer_read(str(ok))
    assert list(ids) == [7] and t[0] == 1e-3 and m[0] == 2.0
    assert np.array_equal(q[0], [1, 2, 3]) and np.array_equal(p[0], [-8, 10, 12])  # '-5' = item dash + 5
    with pytest.raises(nbh.NbhError):          # capacity
        nbh.advancer_read(str(ok), cap=0)
