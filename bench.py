#!/usr/bin/env python
"""bench.py — FP64 pairwise interactions/s of the Hermite acc+jerk sweep
(Hermite.start_bs' pair loop, hermite.ml:154-167 -> nbk_grad_v_dgrad_v_sum) at N = 131072
on 1/2/4/8 B200.

    python bench.py --gpus 1 --steps K --warmup W                 (this arm)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                          (CPU arm, oracle port)

A step = one full acc+jerk sweep: every rank owns a contiguous shard of the targets, all-gathers
the packed (q, v, m) of all bodies with NCCL, and sweeps its targets over all N sources
(strong scaling: N is fixed).  `value` = N(N-1) interactions / step, device-timed with the
bodies resident in HBM; `e2e` = the same sweep through the reference-facing C-ABI call on
pinned HOST buffers (H2D + pack + sweep + D2H inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "direct-nbody_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "FP64 pairwise interactions/sec (Hermite acc+jerk) at N=128k"
UNIT = "interactions/s"
FLOP_PER_INTERACTION = 60        # Nitadori-Makino convention
COUNTED_FLOP = 48                # 6 DADD + 8 DMUL + 17 DFMA(x2) of our instruction mix
FP64_SLOTS = 31                  # FP64-pipe issue slots per interaction (tools/sass_mix.py)
SEED = 20243                     # 20240 + config index 3


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons DURING the timed region: NVML (pynvml, ~5 ms
    period) so that even a 40 ms timed region gets samples; falls back to polling nvidia-smi."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, threading.Event()
        self.sm, self.mx, self.reasons = [], [], set()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for bit, name in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True,
                             timeout=5).stdout.strip()
        f = [x.strip() for x in out.split(",")]
        if len(f) >= 6:
            self.sm.append(float(f[0]))
            self.mx.append(float(f[1]))
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for k in range(4):
                if f[2 + k].lower().startswith("active"):
                    self.reasons.add(names[k])

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.005 if self.nvml else 0.1)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def ncu_traffic_bytes():
    """dram__bytes_read + dram__bytes_write of the acc+jerk sweep kernel per launch, from the
    committed ncu --set full summary (profiles/); None if absent."""
    path = os.path.join(ROOT, "profiles", "r1_ncu_k2_large_shape_summary.txt")
    try:
        tot = 0.0
        for line in open(path):
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if line.startswith(key + " "):
                    unit = line.split("[")[1].split("]")[0]
                    val = float(line.split("=")[1])
                    tot += val * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return tot or None
    except Exception:
        return None


def cpu_sample(n_sample, threads_all=True, n_total=131072):
    """The reference's CPU path (oracle port of Hermite.start_bs' symmetric pair loop) timed on
    a bounded sample: the first n_sample bodies of the same Plummer workload."""
    from oracle import oracle as O
    m, q, p = O.make_plummer(n_total, SEED)
    m, q, p = m[:n_sample].copy(), q[:n_sample].copy(), p[:n_sample].copy()
    t = time.perf_counter()
    O.start_bs_sweep(m, q, p, 0.0, omp=threads_all)
    dt = time.perf_counter() - t
    cores = O.num_threads() if threads_all else 1
    return n_sample * (n_sample - 1) / dt, cores, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The OCaml
    reference cannot be compiled in this image (no OCaml toolchain), so this is the oracle
    port (kind "port"), all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.cpu_sample
    vals = []
    for k in range(args.warmup + args.steps):
        v, cores, dt = cpu_sample(n_sample, True, args.n)
        if k >= args.warmup:
            vals.append((v, dt))
    value = statistics.mean(v for v, _ in vals)
    ms = 1e3 * statistics.mean(d for _, d in vals)
    sample = (f"symmetric i<j acc+jerk sweep over the first {n_sample} bodies of the N={args.n} "
              f"Plummer workload, OpenMP over {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "gpu_launches": 0,
            "config": {"workload": f"Plummer sphere N={args.n}, Hermite acc+jerk sweep "
                                   "(BASELINE.json configs[2])", "n": args.n, "eps": 0.0,
                       "seed": SEED},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print_line(line)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import nbh
    import nbk

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.n
    from nbk.sharded import PeerShardedAccJerk, ShardedAccJerk, make_plan
    plan = make_plan(n, world, rank, align=128)
    t0, t1 = plan.t0, plan.t1
    ctx = nbk.Context(local)
    m, q, p = nbh.make_plummer(n, SEED)  # same seeded bodies on every rank

    # ---- device-resident state: this rank's shard of packed (q, v, m).
    # exchange "peer" (default for N > 1): the rows stay in a region the other ranks map over
    # NVLink (CUDA IPC) and the sweep kernel reads remote source tiles itself - no all-gather;
    # "nccl": in-place all-gather of the packed rows in front of the sweep.
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    exchange = args.exchange if world > 1 else "none"
    sh = None
    if exchange == "peer":
        try:  # raises on EVERY rank if any rank cannot allocate or map a region
            sh = PeerShardedAccJerk(plan, dev, ctx)
        except RuntimeError as e:
            print(f"rank {rank}: {e}; falling back to the NCCL all-gather", file=sys.stderr)
            sh, exchange = None, "nccl (peer exchange unavailable)"
    if sh is None:
        sh = ShardedAccJerk(plan, dev, ctx=ctx, overlap=bool(args.overlap))
    sh.load_local(m, q, p)
    g, dg = sh.g, sh.dg
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        sh.step(0.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.zero_()
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations, outside the per-step events
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    barrier()
    sampler.stop_flag.set()
    gpu_launches = ctx.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    tot_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(tot_ms.item()) / args.steps
    inter = float(n) * float(n - 1)
    value = inter / (ms_per_step * 1e-3)

    # ---- dominant kernel alone (this rank's share), live CUDA events, for the roofline
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    if exchange == "peer":
        barrier()  # every rank's rows are published; the kernel alone needs no flags
        shards = ctx.make_shards(plan.rank, plan.shard, sh.regions.packed(sh.cur))

        def kernel_only():
            ctx.dev_grad_v_dgrad_v_sum_peer(shards, n, 0.0, (t0, t1), (0, n), g.data_ptr(),
                                            dg.data_ptr(), False, sptr)
    else:
        src = sh.packed_all

        def kernel_only():
            ctx.dev_grad_v_dgrad_v_sum(src.data_ptr(), n, 0.0, (t0, t1), (0, n), g.data_ptr(),
                                       dg.data_ptr(), False, sptr)
    for k in range(args.steps):
        flush.zero_()
        kev[k][0].record(stream)
        kernel_only()
        kev[k][1].record(stream)
    torch.cuda.synchronize()
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    k_inter = float(t1 - t0) * float(n - 1)

    # ---- e2e: the reference-facing C-ABI call on pinned host buffers
    hm = torch.from_numpy(m).pin_memory()
    hq = torch.from_numpy(q).pin_memory()
    hp = torch.from_numpy(p).pin_memory()
    nt = t1 - t0
    hg = torch.empty(max(nt, 1), 3, dtype=torch.float64).pin_memory()
    hdg = torch.empty(max(nt, 1), 3, dtype=torch.float64).pin_memory()
    import ctypes as C
    pd = C.POINTER(C.c_double)

    def cptr(t):
        return C.cast(t.data_ptr(), pd)

    def e2e_step():
        rc = ctx._lib.nbk_grad_v_dgrad_v_sum(ctx._h, n, cptr(hm), cptr(hq), cptr(hp), 0.0, t0, t1,
                                             0, n, cptr(hg), cptr(hdg))
        ctx._ck(rc)

    e2e_steps = max(3, min(args.steps, 10))
    e2e_step()
    barrier()
    tw = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()  # synchronous: returns with results in host memory
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - tw) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = inter / float(e2e_s.item())
    h2d = 7 * 8 * n
    d2h = 6 * 8 * nt

    if rank == 0:
        peak_tf, _ = ctx.fp64_peak(20000)
        peak_tf, _ = ctx.fp64_peak(20000)
        achieved_tf = FLOP_PER_INTERACTION * k_inter / (k_ms * 1e-3) / 1e12
        clocks = sampler.summary()
        nominal = ctx.sm_count * 64 * 2 * (clocks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
        try:
            cpu_v, cpu_cores, cpu_dt = cpu_sample(args.cpu_sample, True, n) if world == 1 else (None, None, None)
        except Exception as e:  # the oracle is test infrastructure; its absence must not break the bench
            cpu_v, cpu_cores, cpu_dt = None, None, None
            print(f"cpu_baseline unavailable: {e}", file=sys.stderr)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Plummer sphere N={n}, Hermite acc+jerk sweep "
                                   "(BASELINE.json configs[2])", "n": n, "eps": 0.0, "seed": SEED,
                       "sharding": (
                           "single GPU" if world == 1 else
                           f"targets/{world}, sweep kernel reads remote source tiles from the owning "
                           "GPU over NVLink (peer memory, CUDA IPC); no all-gather step"
                           if exchange == "peer" else
                           f"targets/{world}, NCCL all-gather of packed (q,v,m) per step [{exchange}]"
                           + (", overlapped with the local-source sweep" if args.overlap else "")),
                       "l2": "flushed between timed iterations (256 MiB write)"},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                         "traffic": ncu_traffic_bytes() if world == 1 else None,
                         "traffic_note": "ncu dram bytes per launch (profiles/); algorithmic = "
                                         "64 B x N read once + 48 B x N written",
                         "peak_source": "on-box DFMA microbenchmark (nbk_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "nominal_peak": nominal, "frac_of_nominal": achieved_tf / nominal,
                         "flop_per_interaction": FLOP_PER_INTERACTION,
                         "counted_flop_per_interaction": COUNTED_FLOP,
                         "achieved_counted": COUNTED_FLOP * k_inter / (k_ms * 1e-3) / 1e12,
                         "fp64_issue_slots_per_interaction": FP64_SLOTS,
                         "kernel": "sweep_kernel<OpGradVDGradV" + (", PEER>" if exchange == "peer" else ">"), "kernel_ms": k_ms,
                         "interactions_per_launch": k_inter},
        }
        if cpu_v is not None:
            line["cpu_baseline"] = {
                "value": cpu_v, "unit": UNIT, "cores": cpu_cores, "kind": "port",
                "sample": f"symmetric i<j acc+jerk sweep over the first {args.cpu_sample} bodies of "
                          f"the same Plummer workload ({cpu_dt:.2f} s wall, OpenMP, all host threads)"}
        print_line(line)
    if hasattr(sh, "close"):
        sh.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=131072)
    ap.add_argument("--cpu-sample", type=int, default=65536,
                    help="bodies of the workload the CPU arm sweeps per step (~20 core-seconds)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how a rank gets the other ranks' bodies (see run_ours)")
    ap.add_argument("--overlap", type=int, default=0,
                    help="1: sweep local sources while the all-gather is in flight")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout must carry exactly ONE JSON line: route everything libraries print there (NCCL's
    # version banner, ...) to stderr and give print_line() the real stdout back.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


_REAL_STDOUT = None


def print_line(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    main()
