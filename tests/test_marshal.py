"""CPU: the reference's binary state streams (Marshal.to_channel chan bs [] per checkpoint,
bin/nbody.ml:80, bin/el_mass_segregation.ml:96-108; read back by Analysis.load_states / load_state,
analysis.ml:556-575) as mirrored in host/marshal.cpp.  No OCaml runtime exists in this image, so
the reader is checked against streams ASSEMBLED BY HAND from the runtime's published format
(runtime/caml/intext.h) - including what a real writer produces and ours does not: physical
sharing (the one boxed `nan` of A.make_body_id, advancer.ml:98-116, referenced from every fresh
body), big-endian doubles, every integer width, the 64-bit "big" header - and the writer by
round trips plus an independent re-derivation of its header words.  Not pinned against a real
runtime; oracle/ref_fixtures/gen.ml writes ref/states_adv.bin for that day, and the same gen.ml run
by the interpreter oracle/miniml gives ref_miniml/states_adv.bin (the reference's records through
an independent encoder that DOES share boxed floats), read below."""
import os
import struct

import numpy as np
import pytest

import nbh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _states(n, seed):
    rng = np.random.default_rng(seed)
    st = rng.standard_normal((n, nbh.ADV_STRIDE))
    ids = rng.integers(-40000, 1 << 20, n).astype(np.int32)
    ids[:4] = [0, 63, 64, -1][:min(4, n)]
    return ids, st


def test_round_trip_and_append(tmp_path):
    path = tmp_path / "states.dat"
    written = []
    for k, n in enumerate((5, 1, 300, 0, 7)):
        ids, st = _states(n, k)
        if n > 2:
            st[1, 8:] = np.nan   # a fresh body: NaN after p
            st[2, 0] = np.inf
        nbh.marshal_write_state(path, ids, st, append=k > 0)
        written.append((ids, st))
    assert nbh.marshal_count_states(path) == 5
    for k, (ids, st) in enumerate(written):
        kind, rid, rst = nbh.marshal_read_state(path, k)
        assert (kind == 17 or len(ids) == 0) and np.array_equal(rid, ids)
        assert rst.shape == st.shape and np.array_equal(rst.view(np.uint64), st.view(np.uint64))
    assert len(nbh.marshal_load_states(path)) == 5
    with pytest.raises(nbh.NbhError):  # load_state past the end raises End_of_file
        nbh.marshal_read_state(path, 5)
    # a truncated last value is ignored, like load_states' `with _ ->`
    data = path.read_bytes()
    (tmp_path / "cut.dat").write_bytes(data[:-10])
    assert nbh.marshal_count_states(tmp_path / "cut.dat") == 4


def test_writer_header_words_follow_the_runtime_rules(tmp_path):
    """magic, data length, object count and the 32-/64-bit heap sizes, re-derived here from the
    format definition: array block 1+n words; record 1+17; boxed float 1+1 (64-bit) / 1+2
    (32-bit); float array of 3: 1+3 / 1+6.  One object per block, boxed float and float array."""
    n = 9
    ids, st = _states(n, 3)
    path = tmp_path / "s.dat"
    nbh.marshal_write_state(path, ids, st, append=False)
    data = path.read_bytes()
    magic, dlen, nobj, w32, w64 = struct.unpack(">5I", data[:20])
    assert magic == 0x8495A6BE and dlen == len(data) - 20
    floats, vecs = 5, 10  # t m h hmax t0; q p dVdq0-2 p0 q0 q1 q2 cs
    assert nobj == 1 + n * (1 + floats + vecs)
    assert w64 == (1 + n) + n * ((1 + 17) + floats * 2 + vecs * 4)
    assert w32 == (1 + n) + n * ((1 + 17) + floats * 3 + vecs * 7)
    assert data[20] == 0x08 and struct.unpack(">I", data[21:25])[0] == (n << 10)  # array of 9: BLOCK32
    assert data[25] == 0x08 and struct.unpack(">I", data[26:30])[0] == (17 << 10)  # first record
    assert data[30] == 0x40  # id 0 as a small int


def _hand_stream(big_header=False, big_endian=False):
    """Two Advancer.A bodies as ocamlopt would marshal FRESH bodies: t0, h, hmax point at ONE
    boxed nan (first occurrence written, later ones CODE_SHARED8), so do the NaN vectors."""
    dcode, acode, fmt = (0x0B, 0x0D, ">d") if big_endian else (0x0C, 0x0E, "<d")
    out, nobj = bytearray(), 0

    def boxed(x):
        nonlocal nobj
        out.append(dcode)
        out.extend(struct.pack(fmt, x))
        nobj += 1
        return nobj

    def vec(v):
        nonlocal nobj
        out.append(acode)
        out.append(3)
        for x in v:
            out.extend(struct.pack(fmt, x))
        nobj += 1
        return nobj

    def shared(obj):
        out.extend([0x04, nobj + 1 - obj]) if nobj + 1 - obj < 256 else out.extend(
            [0x05] + list(struct.pack(">H", nobj + 1 - obj)))

    out.append(0x80 | 0 | (2 << 4))  # array of 2: small block
    nobj += 1
    bodies = [(7, 0.25, 0.5, (1.0, 2.0, 3.0), (-1.0, -2.0, -3.0)),
              (-300, 0.25, 2.0, (4.0, 5.0, 6.0), (0.5, 0.25, 0.125))]
    nan_obj = nanvec_obj = None
    for k, (bid, t, m, q, p) in enumerate(bodies):
        out.append(0x08)
        out.extend(struct.pack(">I", 17 << 10))
        nobj += 1
        if 0 <= bid < 64:
            out.append(0x40 | bid)
        else:
            out.append(0x01)
            out.extend(struct.pack(">h", bid))
        boxed(t)
        boxed(m)
        vec(q)
        vec(p)
        for _ in range(3):  # h, hmax, t0 : the same nan
            if nan_obj is None:
                nan_obj = boxed(float("nan"))
            else:
                shared(nan_obj)
        for _ in range(8):  # dVdq0-2, p0, q0, q1, q2, cs: Array.make 3 nan, one per field here...
            if k == 0 or nanvec_obj is None:
                nanvec_obj = vec((float("nan"),) * 3)
            else:           # ...and, to exercise it, body 1 shares body 0's last vector
                shared(nanvec_obj)
        out.append(0x40)    # output_counter = 0
    if big_header:
        hdr = struct.pack(">II3Q", 0x8495A6BF, 0, len(out), nobj, 0)
    else:
        hdr = struct.pack(">5I", 0x8495A6BE, len(out), nobj, 0, 0)
    return hdr + bytes(out), bodies


@pytest.mark.parametrize("big_header,big_endian", [(False, False), (True, False), (False, True)])
def test_reader_on_hand_assembled_streams_with_sharing(tmp_path, big_header, big_endian):
    data, bodies = _hand_stream(big_header, big_endian)
    path = tmp_path / "hand.dat"
    path.write_bytes(data + data)  # two states back to back
    assert nbh.marshal_count_states(path) == 2
    for k in range(2):
        kind, ids, st = nbh.marshal_read_state(path, k)
        assert kind == 17 and list(ids) == [b[0] for b in bodies]
        for r, (bid, t, m, q, p) in enumerate(bodies):
            assert st[r, 0] == t and st[r, 1] == m
            assert tuple(st[r, 2:5]) == q and tuple(st[r, 5:8]) == p
            assert np.isnan(st[r, 8:]).all()


def test_reader_accepts_leapfrog_records_and_rejects_garbage(tmp_path):
    out = bytearray([0x80 | (1 << 4)])  # array of 1
    out.append(0x80 | (6 << 4))         # Leapfrog.b: id t dt_max m q v
    out.append(0x40 | 5)
    for x in (1.5, 0.01, 0.125):
        out.append(0x0C)
        out.extend(struct.pack("<d", x))
    for v in ((1.0, 2.0, 3.0), (4.0, 5.0, 6.0)):
        out.extend([0x0E, 3])
        for x in v:
            out.extend(struct.pack("<d", x))
    path = tmp_path / "lf.dat"
    path.write_bytes(struct.pack(">5I", 0x8495A6BE, len(out), 6, 0, 0) + bytes(out))
    kind, ids, st = nbh.marshal_read_state(path, 0)
    assert kind == 6 and list(ids) == [5]
    assert st[0, 0] == 1.5 and st[0, 8] == 0.01 and st[0, 1] == 0.125
    assert tuple(st[0, 2:5]) == (1.0, 2.0, 3.0) and tuple(st[0, 5:8]) == (4.0, 5.0, 6.0)
    bad = tmp_path / "bad.dat"
    bad.write_bytes(b"--- !Particle\nid: 0\n" * 3)
    with pytest.raises(nbh.NbhError):
        nbh.marshal_count_states(bad)
    bad.write_bytes(struct.pack(">5I", 0x8495A6BE, 3, 0, 0, 0) + bytes([0x04, 0x09, 0x00]))
    with pytest.raises(nbh.NbhError):  # shared reference with no object before it
        nbh.marshal_read_state(bad, 0)


@pytest.mark.parametrize("where", ["ref", "ref_miniml"])
def test_reference_written_stream_if_present(where):
    """ref/states_adv.bin comes from the REAL Marshal module (oracle/ref_fixtures/gen.ml compiled
    with ocamlfind; absent until someone does that).  ref_miniml/states_adv.bin is the same
    gen.ml interpreted by oracle/miniml over the unmodified Advancer.A: the body records are the
    reference's, the encoder is miniml's (it shares boxed floats by identity, so the stream is
    full of SHARED back-references - a different byte string from ocamlopt's, the same value)."""
    kit = os.path.join(ROOT, "oracle", "ref_fixtures")
    path = os.path.join(kit, where, "states_adv.bin")
    if not os.path.exists(path):
        pytest.skip("no stream written by a real OCaml runtime (oracle/ref_fixtures/ref/states_adv.bin)")
    import sys
    sys.path.insert(0, kit)
    import make_inputs
    m, q, p = make_inputs.read_bodies(os.path.join(kit, "inputs", "plummer_n16_std_seed20241.txt"))
    assert nbh.marshal_count_states(path) == 2
    kind, ids, st = nbh.marshal_read_state(path, 0)   # the fresh bodies
    assert kind == 17 and np.array_equal(st[:, 1], m) and np.array_equal(st[:, 2:5], q)
    assert np.array_equal(st[:, 5:8], p) and np.isnan(st[:, 8:]).all()
    kind, ids, st = nbh.marshal_read_state(path, 1)   # after A.advance bs 0.125 1e-3
    ref = {}
    for ln in open(os.path.join(kit, where, "traj.hex")):
        key, i, j, k, bits = ln.split()
        ref[(key, int(i), int(k))] = struct.unpack("<d", struct.pack("<Q", int(bits, 16)))[0]
    for r in range(len(ids)):
        assert st[r, 0] == ref[("adv_t", r, 0)] and st[r, 9] == ref[("adv_hmax", r, 0)]
        for k in range(3):
            assert st[r, 2 + k] == ref[("adv_q", r, k)] and st[r, 5 + k] == ref[("adv_p", r, k)]
