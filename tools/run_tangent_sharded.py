#!/usr/bin/env python
"""BASELINE configs[4]: DFloat pushforward tangent (variational) forces, N = 32768, K directions,
sharded over the GPUs of one box.  Launch with torchrun (or plain python for 1 GPU):
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_tangent_sharded.py
Prints one JSON line on rank 0: tangent interactions/s = K N (N-1) / step."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402
from nbk.sharded import PeerShardedTangent, ShardedTangent, make_plan  # noqa: E402

n = int(os.environ.get("NBK_N", "32768"))
K = int(os.environ.get("NBK_K", "6"))
steps = int(os.environ.get("NBK_STEPS", "10"))
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = nbk.Context(local)
m, q, p = nbh.make_plummer(n, 20245)
rng = np.random.default_rng(20245)
dq = rng.standard_normal((K, n, 3))
dp = rng.standard_normal((K, n, 3)) * m[None, :, None]
predicted = os.environ.get("NBK_PREDICTED", "0") == "1"
if predicted:
    # the dual-number Hermite step's shape: bodies at individual times
    # with tangents of their whole state, predicted to a dual time on the device, peer exchange
    g0, dg0 = ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)
    f, df = -g0, -dg0
    t = rng.uniform(0.0, 1e-3, n)
    t_tan = rng.standard_normal((K, n)) * 1e-3
    f_tan = rng.standard_normal((K, n, 3)) * np.abs(f).mean()
    df_tan = rng.standard_normal((K, n, 3)) * np.abs(df).mean()
    nt, nt_tan = 2e-3, rng.standard_normal(K) * 1e-3
    plan = make_plan(n, world, rank, align=128)
    sh = PeerShardedTangent(plan, K, dev, ctx)
    sh.load_local_predicted(t, m, q, p, f, df, t_tan, dq, dp, f_tan, df_tan, nt, nt_tan)
else:
    plan = make_plan(n, world, rank)
    sh = ShardedTangent(plan, K, dev, ctx)
    sh.load_local(m, q, p, dq, dp)
for _ in range(3):
    sh.step(0.0)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    gt, dgt = sh.step(0.0)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# parity of this rank's first 8 targets against the single-call host API on rank 0's own GPU
ok = True
if rank == 0:
    if predicted:
        ref = ctx.grad_v_dgrad_v_sum_predicted_tangent(t, m, q, p, f, df, t_tan, dq, dp, f_tan,
                                                       df_tan, nt, nt_tan, 0.0, targets=(0, 8))
    else:
        ref = ctx.grad_v_dgrad_v_sum_tangent(m, q, p, dq, dp, 0.0, targets=(0, 8))
    err = float(np.max(np.linalg.norm(gt[:, :8].cpu().numpy() - ref[2], axis=-1)
                       / np.linalg.norm(ref[2], axis=-1)))
    err = max(err, float(np.max(np.linalg.norm(dgt[:, :8].cpu().numpy() - ref[3], axis=-1)
                                / np.linalg.norm(ref[3], axis=-1))))
    print(json.dumps({"config": 4, "what": "sharded tangent acc+jerk sums"
                      + (" on bodies predicted over duals (DFloat_hermite.advance_body shape, "
                         "nbk_dev_predict_pack_tangent + peer-read tangent sweep)" if predicted else ""),
                      "n": n, "K": K,
                      "n_gpus": world, "ms_per_step": float(ms.item()),
                      "tangent_interactions_per_s": K * n * (n - 1.0) / (float(ms.item()) * 1e-3),
                      "max_rel_diff_vs_single_gpu_call": err}))
if hasattr(sh, "close"):
    sh.close()
ctx.close()
if world > 1:
    dist.destroy_process_group()
