"""Target-sharded sweeps across the GPUs of one box: one process per GPU,
torch.distributed for the plumbing.

The path shards by TARGETS: rank r owns the contiguous slice [t0, t1) of the bodies and needs
every body as a source, so each sweep has exactly one exchange step - an all-gather of the
packed (q, v, m) rows, 64 B per body - followed by the rank's targets x all sources sweep.
Outputs stay sharded.  With overlap=True the all-gather runs asynchronously while the rank
sweeps its own slice as sources, and the remote sources are swept (accumulating) once it lands.

The collective and the sweep are injectable so the plumbing is testable on CPU with gloo
(tests/test_sharded_gloo.py); the defaults are NCCL and the CUDA kernel - no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass

PACKED = 8  # doubles per packed body


@dataclass(frozen=True)
class ShardPlan:
    n: int        # bodies
    world: int    # ranks
    rank: int
    shard: int    # rows per rank in the gathered array (last rank's tail rows are padding)
    t0: int       # this rank's targets [t0, t1)
    t1: int

    @property
    def npad(self) -> int:
        return self.shard * self.world

    @property
    def count(self) -> int:
        return self.t1 - self.t0


def make_plan(n: int, world: int, rank: int) -> ShardPlan:
    if n < 0 or world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad shard plan n={n} world={world} rank={rank}")
    shard = (n + world - 1) // world if n else 0
    t0, t1 = min(rank * shard, n), min((rank + 1) * shard, n)
    return ShardPlan(n, world, rank, shard, t0, t1)


class ShardedAccJerk:
    """One rank of the sharded Hermite acc+jerk sweep (hermite.ml:154-167 over all bodies)."""

    def __init__(self, plan: ShardPlan, device, ctx=None, group=None, overlap=False,
                 all_gather=None, sweep=None):
        import torch
        import torch.distributed as dist
        self.plan, self.device, self.ctx, self.group, self.overlap = plan, device, ctx, group, overlap
        self.torch = torch
        # row i of packed_all is body i; this rank's rows are a view, so the gather is in place
        self.packed_all = torch.zeros(max(plan.npad, 1), PACKED, dtype=torch.float64, device=device)
        lo = plan.rank * plan.shard
        self.packed_loc = self.packed_all[lo:lo + plan.shard]
        self.g = torch.zeros(max(plan.count, 1), 3, dtype=torch.float64, device=device)
        self.dg = torch.zeros_like(self.g)
        if all_gather is None:
            def all_gather(out, inp, async_op=False):
                return dist.all_gather_into_tensor(out, inp, group=group, async_op=async_op)
        self._all_gather = all_gather
        self._sweep = sweep if sweep is not None else self._cuda_sweep
        if sweep is None and ctx is None:
            raise RuntimeError("ShardedAccJerk needs an nbk.Context (no CPU fallback)")

    # ---- default sweep: the CUDA kernel through the C ABI, on torch's current stream
    def _cuda_sweep(self, packed, n, eps2, targets, sources, g, dg, accumulate):
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        self.ctx.dev_grad_v_dgrad_v_sum(packed.data_ptr(), n, eps2, targets, sources,
                                        g.data_ptr(), dg.data_ptr(), accumulate, stream)

    def load_local(self, m, q, p):
        """Pack this rank's bodies (host arrays of the FULL system, or device tensors of the
        local slice) into its rows of the gathered array."""
        torch, pl = self.torch, self.plan
        if pl.count == 0:
            return
        sl = slice(pl.t0, pl.t1)
        d_m = torch.as_tensor(m[sl], dtype=torch.float64).to(self.device).contiguous()
        d_q = torch.as_tensor(q[sl], dtype=torch.float64).to(self.device).contiguous()
        d_p = torch.as_tensor(p[sl], dtype=torch.float64).to(self.device).contiguous()
        if self.ctx is not None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self.ctx.dev_pack(pl.count, d_m.data_ptr(), d_q.data_ptr(), d_p.data_ptr(), False,
                              self.packed_loc.data_ptr(), stream)
            torch.cuda.current_stream(self.device).synchronize()
        else:  # CPU plumbing tests: same layout {x,y,z,m,vx,vy,vz,0}
            rows = self.packed_loc[:pl.count]
            rows[:, 0:3] = d_q
            rows[:, 3] = d_m
            rows[:, 4:7] = d_p / d_m[:, None]
            rows[:, 7] = 0.0

    def step(self, eps2: float = 0.0):
        """All-gather, then sweep this rank's targets over all sources.  Returns the device
        tensors (gsum, dgsum) for targets [t0, t1)."""
        pl = self.plan
        if pl.world == 1:
            if pl.count:
                self._sweep(self.packed_all, pl.n, eps2, (pl.t0, pl.t1), (0, pl.n), self.g,
                            self.dg, False)
            return self.g, self.dg
        if not self.overlap or pl.count == 0:
            self._all_gather(self.packed_all, self.packed_loc)
            if pl.count:
                self._sweep(self.packed_all, pl.n, eps2, (pl.t0, pl.t1), (0, pl.n), self.g,
                            self.dg, False)
            return self.g, self.dg
        # overlap: own slice as sources while the gather is in flight, then the rest
        work = self._all_gather(self.packed_all, self.packed_loc, async_op=True)
        self._sweep(self.packed_all, pl.n, eps2, (pl.t0, pl.t1), (pl.t0, pl.t1), self.g, self.dg,
                    False)
        if work is not None:
            work.wait()
        if pl.t0 > 0:
            self._sweep(self.packed_all, pl.n, eps2, (pl.t0, pl.t1), (0, pl.t0), self.g, self.dg,
                        True)
        if pl.t1 < pl.n:
            self._sweep(self.packed_all, pl.n, eps2, (pl.t0, pl.t1), (pl.t1, pl.n), self.g,
                        self.dg, True)
        return self.g, self.dg


class ShardedTangent:
    """One rank of the sharded tangent (variational) acc+jerk sums for K directions
    (dFloat_hermite.ml:177-186 over all bodies; BASELINE configs[4]).  Exchange step: an in-place
    all-gather of the packed bodies and of the K packed direction arrays."""

    def __init__(self, plan: ShardPlan, K: int, device, ctx, group=None):
        import torch
        import torch.distributed as dist
        self.plan, self.K, self.device, self.ctx, self.group = plan, K, device, ctx, group
        self.torch, self.dist = torch, dist
        npad = max(plan.npad, 1)
        self.packed_all = torch.zeros(npad, PACKED, dtype=torch.float64, device=device)
        # direction-major [K][npad][8] so that the kernel's aux stride is npad*8
        self.aux_all = torch.zeros(K, npad, PACKED, dtype=torch.float64, device=device)
        lo = plan.rank * plan.shard
        self.lo = lo
        self.gt = torch.zeros(K, max(plan.count, 1), 3, dtype=torch.float64, device=device)
        self.dgt = torch.zeros_like(self.gt)

    def load_local(self, m, q, p, dq, dp):
        """m,q,p: full-system host arrays; dq, dp: [K][n][3] host arrays."""
        torch, pl = self.torch, self.plan
        if pl.count == 0:
            return
        sl = slice(pl.t0, pl.t1)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        d_m = torch.as_tensor(m[sl]).to(self.device).contiguous()
        d_q = torch.as_tensor(q[sl]).to(self.device).contiguous()
        d_p = torch.as_tensor(p[sl]).to(self.device).contiguous()
        self.ctx.dev_pack(pl.count, d_m.data_ptr(), d_q.data_ptr(), d_p.data_ptr(), False,
                          self.packed_all[self.lo:].data_ptr(), stream)
        keep = [d_m, d_q, d_p]
        for k in range(self.K):
            d_dq = torch.as_tensor(dq[k][sl]).to(self.device).contiguous()
            d_dp = torch.as_tensor(dp[k][sl]).to(self.device).contiguous()
            self.ctx.dev_pack_aux(pl.count, d_m.data_ptr(), d_dq.data_ptr(), d_dp.data_ptr(),
                                  self.aux_all[k, self.lo:].data_ptr(), stream)
            keep += [d_dq, d_dp]
        torch.cuda.current_stream(self.device).synchronize()

    def step(self, eps2: float = 0.0):
        pl, dist = self.plan, self.dist
        if pl.world > 1:
            dist.all_gather_into_tensor(self.packed_all, self.packed_all[self.lo:self.lo + pl.shard],
                                        group=self.group)
            for k in range(self.K):
                dist.all_gather_into_tensor(self.aux_all[k],
                                            self.aux_all[k, self.lo:self.lo + pl.shard],
                                            group=self.group)
        if pl.count:
            stream = self.torch.cuda.current_stream(self.device).cuda_stream
            self.ctx.dev_grad_v_dgrad_v_sum_tangent(
                self.packed_all.data_ptr(), self.aux_all.data_ptr(), pl.npad, self.K, eps2,
                (pl.t0, pl.t1), (0, pl.n), self.gt.data_ptr(), self.dgt.data_ptr(), stream)
        return self.gt, self.dgt
