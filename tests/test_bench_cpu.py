"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port timed on host cores)
prints exactly one JSON line with the contract's keys; the GPU arm refuses to run without a GPU
instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT,
                          capture_output=True, text=True, timeout=300, env=e)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "2048",
             "--n", "8192")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "interactions/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # hygiene: thread count set by the arm itself, single-thread rate beside it,
    # the stated N measured at least once, config identical to the GPU arm's
    assert d["cpu_baseline"]["full_n"]["measured"] is True and d["cpu_baseline"]["full_n"]["n"] == 8192
    assert d["cpu_baseline"]["single_thread"]["cores"] == 1
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.make_config(8192, 1)


def test_reference_arm_ignores_omp_num_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must still use every host core."""
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "2048",
             "--n", "4096", env={"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip())
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_parity_ranges_cover_first_and_last_tile():
    sys.path.insert(0, ROOT)
    import bench
    for t0, t1 in ((0, 131072), (16384, 32768), (114688, 131072), (0, 300), (5, 100), (7, 7)):
        rs = bench.parity_ranges(t0, t1)
        tot = sum(b - a for a, b in rs)
        assert tot == min(256, t1 - t0)
        if rs:
            assert rs[0][0] == t0 and rs[-1][1] == t1
            assert all(t0 <= a < b <= t1 for a, b in rs)
            assert all(x[1] <= y[0] for x, y in zip(rs, rs[1:]))


def test_reference_arm_other_ranks_print_nothing():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run("--steps", "1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_config_states_the_pair_convention_and_both_arms_agree():
    """`--pairs` is part of the stated workload: both arms print the same `config`, and it says
    whether every unordered pair is evaluated once (csrc/sym.cuh) or every ordered pair."""
    sys.path.insert(0, ROOT)
    import bench
    for world in (1, 2, 8):
        once = bench.make_config(131072, world, "peer", 0, "once")
        ordered = bench.make_config(131072, world, "peer", 0, "ordered")
        assert "once" in once["pairs"] and "ordered" in ordered["pairs"]
        assert once == bench.make_config(131072, world)  # the default both arms use
        assert set(once) == set(ordered)
        if world > 1:
            assert "block pairs" in once["sharding"] and "remote source tiles" in ordered["sharding"]
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "2048",
             "--n", "4096", "--pairs", "ordered")
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout.strip())["config"] == bench.make_config(4096, 1, "peer", 0, "ordered")


def test_roofline_constants_and_committed_traffic():
    """FP64-pipe slots per ordered interaction by kernel (19 = 38 per pair evaluation / 2), and
    the ncu traffic figure is read from the SWEEP kernel's record of the committed summary."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.FP64_SLOTS == {"ordered": 31.0, "once": 19.0}
    assert bench.COUNTED_FLOP["once"] == (6 + 9 + 2 * 23) / 2
    assert bench.COUNTED_FLOP["ordered"] == 6 + 8 + 2 * 17
    t_once, f_once = bench.ncu_traffic_bytes("once")
    t_ord, f_ord = bench.ncu_traffic_bytes("ordered")
    assert f_once == "profiles/r2_ncu_sym_summary.txt" and 7.0e8 < t_once < 8.5e8
    assert f_ord.endswith("k2_large_shape_summary.txt") and 8.0e6 < t_ord < 9.0e6
