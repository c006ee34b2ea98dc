"""The bookkeeping of the pair-once sweep (csrc/sym.cuh), restated in Python and checked
exhaustively on CPU: item numbering, slot ownership, the warp-skewed walk over the source
groups between CTA barriers, the lane rotation of the travelling j-side sums, the doubled-group
shared-memory index, the per-rank item ranges and the model that picks the kernel.  The CUDA
side is tests/test_sym_gpu.py."""
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 1024


def item_of(i, j, nb):  # sym_item_of
    return i * nb - i * (i - 1) // 2 + (j - i)


def decode(k, nb):  # sym_decode: closed form + the two correcting loops
    b = 2.0 * nb + 1.0
    i = int((b - math.sqrt(b * b - 8.0 * k)) * 0.5)
    i = max(0, min(nb - 1, i))
    while i > 0 and item_of(i, i, nb) > k:
        i -= 1
    while i + 1 < nb and item_of(i + 1, i + 1, nb) <= k:
        i += 1
    return i, i + (k - item_of(i, i, nb))


@pytest.mark.parametrize("nb", [1, 2, 3, 7, 16, 128, 129, 683])
def test_item_numbering_is_the_row_major_upper_triangle(nb):
    from nbk.sharded import sym_item_blocks
    items = nb * (nb + 1) // 2
    pairs = list(sym_item_blocks(nb * B, 0, items))
    assert len(pairs) == items == len(set(pairs))
    step = max(1, items // 4000)
    for k in list(range(0, items, step)) + [items - 1]:
        assert decode(k, nb) == pairs[k]
        i, j = pairs[k]
        assert 0 <= i <= j < nb and item_of(i, j, nb) == k


@pytest.mark.parametrize("nb", [1, 2, 5, 12])
def test_every_slot_of_every_block_has_exactly_one_writer(nb):
    """Body block Bk has one slot per partner block S; item (I, J) writes slot J of block I
    (targets) and, off the diagonal, slot I of block J (sources)."""
    writers = np.zeros((nb, nb), dtype=int)  # [block][slot]
    for k in range(nb * (nb + 1) // 2):
        i, j = decode(k, nb)
        writers[i, j] += 1
        if i != j:
            writers[j, i] += 1
    assert np.all(writers == 1)
    # and the reduction's rule for a partial item range: slot s of block b is filled iff item
    # (min(b,s), max(b,s)) is in the range
    items = nb * (nb + 1) // 2
    k0, k1 = (1, items - 1) if items >= 3 else (0, items)
    filled = np.zeros((nb, nb), dtype=bool)
    for k in range(k0, k1):
        i, j = decode(k, nb)
        filled[i, j] = filled[j, i] = True
    for b in range(nb):
        for s in range(nb):
            assert filled[b, s] == (k0 <= item_of(min(b, s), max(b, s), nb) < k1)


def test_warps_never_share_a_group_between_barriers():
    """8 warps, 32 groups of 32 sources: in super-step g (one CTA barrier each) warp w walks
    groups ((w + g) & 7) * 4 + {0..3}.  Distinct warps touch distinct groups inside a
    super-step, and every warp visits every group exactly once per item."""
    seen = [set() for _ in range(8)]
    for g in range(8):
        touched = {}
        for w in range(8):
            for s4 in range(4):
                grp = (((w + g) & 7) << 2) + s4
                assert touched.setdefault(grp, w) == w
                assert grp not in seen[w]
                seen[w].add(grp)
    assert all(s == set(range(32)) for s in seen)


def test_travelling_sums_arrive_home_after_32_rounds():
    """Round r: lane l adds the contribution of source (l + r) mod 32 to the registers it
    holds, then the registers move one lane down (lane l reads lane l + 1).  After 32 rounds
    lane l holds the complete sum of source l over all lanes' targets."""
    rng = np.random.default_rng(0)
    contrib = rng.standard_normal((32, 32))  # [lane (its targets)][source]
    ja = np.zeros(32)
    for r in range(32):
        for lane in range(32):
            ja[lane] += contrib[lane, (lane + r) % 32]
        ja = ja[(np.arange(32) + 1) % 32]
    assert np.allclose(ja, contrib.sum(axis=0), rtol=0, atol=1e-13)


def test_doubled_group_entry_index_needs_no_wrap():
    """Source jl of the block sits at entries e and e + 32, e = (jl >> 5 << 6) + (jl & 31);
    lane l in round r of group grp reads entry 2 * grp + l + r = source grp + (l + r) mod 32."""
    where = {}
    for jl in range(B):
        e = ((jl >> 5) << 6) + (jl & 31)
        where[e] = where[e + 32] = jl
    assert sorted(where) == list(range(2 * B))
    for grp in range(0, B, 32):
        for lane in range(32):
            for r in range(32):
                assert where[2 * grp + lane + r] == grp + (lane + r) % 32


@pytest.mark.parametrize("items,world", [(8256, 8), (8256, 3), (10, 4), (3, 8), (0, 2)])
def test_rank_item_ranges_partition_the_items(items, world):
    from nbk.sharded import sym_item_range
    rs = [sym_item_range(items, world, r) for r in range(world)]
    assert rs[0][0] == 0 and rs[-1][1] == items
    assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    sizes = [b - a for a, b in rs]
    assert max(sizes) - min(sizes) <= 1


def test_kernel_constants_match_this_restatement():
    """The numbers this file restates are the ones csrc/sym.cuh compiles with."""
    src = open(os.path.join(ROOT, "direct-nbody_b200", "csrc", "sym.cuh")).read()
    assert re.search(r"SYM_NT = 256;", src) and re.search(r"SYM_IPT = 4;", src)
    assert "((((warp + g) & 7) << 2) + s4) << 5" in src
    assert "((jl >> 5) << 6) + (jl & 31)" in src
    assert "(lane + 1) & 31" in src
    from nbk.sharded import SYM_BLOCK
    assert SYM_BLOCK == B
