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


def make_plan(n: int, world: int, rank: int, align: int = 1) -> ShardPlan:
    """Contiguous equal shards; `align` rounds the shard up to whole source tiles (128 for the
    peer-memory exchange, whose tiles must not straddle two owners)."""
    if n < 0 or world < 1 or not 0 <= rank < world or align < 1:
        raise ValueError(f"bad shard plan n={n} world={world} rank={rank} align={align}")
    shard = (n + world - 1) // world if n else 0
    shard = (shard + align - 1) // align * align
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

    def load_local_device(self, d_m, d_q, d_p):
        """As load_local for device tensors of this rank's slice; enqueued, not waited for."""
        pl = self.plan
        if pl.count == 0:
            return
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        self.ctx.dev_pack(pl.count, d_m.data_ptr(), d_q.data_ptr(), d_p.data_ptr(), False,
                          self.packed_loc.data_ptr(), stream)

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


# ---------------------------------------------------------------------------------------------
# Peer-memory exchange: the all-gather fused into the sweep (include/nbk.h, "peer-memory
# exchange").  Every rank keeps its packed rows in a region the other ranks map over NVLink
# (CUDA IPC); the sweep kernel bulk-copies each source tile straight out of its owner's memory
# behind the TMA ring that stages local tiles, so there is no exchange step left in a sweep.
# ---------------------------------------------------------------------------------------------
PEER_TILE = 128        # sources per tile: shards are whole tiles
PEER_FLAG_BYTES = 1024  # ready[8] flags, padded


PEER_FLAGS2_OFFSET = 512  # a second flag array inside the flag block (the pair-once sums)


def peer_region_layout(shard_rows: int, K: int = 0, nbuf: int = 2, extra_bytes: int = 0):
    """Byte offsets inside one rank's region: flags, then nbuf x {packed rows, K direction
    arrays}, then `extra_bytes` more (at total - extra_bytes, 128-byte aligned).  Returns
    (total_bytes, [packed offset per buffer], [aux offset per buffer])."""
    rows_bytes = shard_rows * PACKED * 8
    buf_bytes = rows_bytes * (1 + K)
    packed_off = [PEER_FLAG_BYTES + b * buf_bytes for b in range(nbuf)]
    aux_off = [o + rows_bytes for o in packed_off]
    return PEER_FLAG_BYTES + nbuf * buf_bytes + extra_bytes, packed_off, aux_off


def next_row_buffer(cur: int, stepped_since_load: bool) -> int:
    """Which of the two row buffers a load may overwrite.  The buffer in use (`cur`) may be read
    by a peer's sweep of the latest epoch for as long as this rank has not completed a LATER
    sweep, so it is never the target once a step has published it; the other buffer was last
    read in sweeps that ended before the peers published the epoch this rank's last sweep
    acquired, so it is free - however many steps ran since the last load (the toggle follows
    the loads, not the epoch parity: load, step, step, load must not reuse the live buffer).
    Two loads without a step in between re-pack the same, still unpublished, buffer."""
    return 1 - cur if stepped_since_load else cur


class PeerRegions:
    """This rank's region plus the mapped regions of every other rank."""

    def __init__(self, plan: ShardPlan, ctx, K: int = 0, group=None, exchange=None,
                 rows_per_buffer=None, extra_bytes=0):
        if plan.shard % PEER_TILE:
            raise ValueError("peer exchange needs make_plan(..., align=128)")
        self.plan, self.ctx, self.K = plan, ctx, K
        # rows_per_buffer: a row buffer that holds more than the rank's own shard (the gathered
        # array of the pair-once sweep, filled by every rank's remote stores)
        self.rows = plan.shard if rows_per_buffer is None else rows_per_buffer
        self.total, self.packed_off, self.aux_off = peer_region_layout(self.rows, K,
                                                                       extra_bytes=extra_bytes)
        self.extra_off = self.total - extra_bytes
        self.base = [0] * plan.world
        self._opened = []
        self.group = group
        if exchange is None:
            def exchange(obj):
                import torch.distributed as dist
                out = [None] * plan.world
                dist.all_gather_object(out, obj, group=group)
                return out
        # Every rank takes part in both exchanges whatever happened to it locally, so that a
        # rank whose allocation or mapping failed cannot leave the others waiting; the outcome
        # (all mapped / RuntimeError) is the same on every rank.
        handle, err = None, None
        try:
            self.base[plan.rank], handle = ctx.peer_alloc(self.total, want_handle=plan.world > 1)
        except Exception as e:  # noqa: BLE001 - reported to every rank below
            err = f"rank {plan.rank}: {e}"
        if plan.world > 1:
            handles = exchange(handle)
            if err is None and all(h is not None for h in handles):
                try:
                    for r in range(plan.world):
                        if r != plan.rank:
                            self.base[r] = ctx.peer_open(handles[r])
                            self._opened.append(self.base[r])
                except Exception as e:  # noqa: BLE001
                    err = f"rank {plan.rank}: {e}"
            errs = [e for e in exchange(err) if e is not None]
            if errs or any(h is None for h in handles):
                self.close(barrier=lambda: exchange(None))
                raise RuntimeError("peer-memory exchange unavailable: " + "; ".join(errs or ["no handle"]))
        elif err is not None:
            raise RuntimeError(err)

    def packed(self, buf):      # rank -> address of its packed rows in buffer `buf`
        return [b + self.packed_off[buf] for b in self.base]

    def aux(self, buf):
        return [b + self.aux_off[buf] for b in self.base]

    @property
    def ready_local(self):
        return self.base[self.plan.rank]

    def close(self, barrier=None):
        for ptr in self._opened:
            self.ctx.peer_close(ptr)
        self._opened = []
        if barrier is not None:
            barrier()  # nobody frees a region a peer still maps
        if self.base[self.plan.rank]:
            self.ctx.peer_free(self.base[self.plan.rank])
            self.base[self.plan.rank] = 0


class _PeerSharded:
    """Common part of the peer-exchange drivers: this rank's rows in the shared region, the
    publish step and the shard table of the current row buffer."""

    def __init__(self, plan: ShardPlan, device, ctx, group=None, exchange=None, K=0,
                 rows_per_buffer=None, extra_bytes=0):
        import torch
        self.torch, self.plan, self.device, self.ctx, self.group = torch, plan, device, ctx, group
        self.regions = PeerRegions(plan, ctx, K, group, exchange, rows_per_buffer, extra_bytes)
        self.epoch, self.cur = 0, 0
        self._stepped_since_load = True  # the first load_local takes the other buffer

    def _stream(self):
        if getattr(self.device, "type", "cuda") != "cuda":  # CPU plumbing tests
            return 0
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def load_local(self, m, q, vel, is_velocity=False):
        """Pack this rank's bodies into the row buffer the NEXT step reads (the one no peer is
        reading: buffers alternate, see include/nbk.h).  vel = momenta, or velocities."""
        torch, pl = self.torch, self.plan
        self.cur = next_row_buffer(self.cur, self._stepped_since_load)
        self._stepped_since_load = False
        if pl.count == 0:
            return
        sl = slice(pl.t0, pl.t1)
        d_m = torch.as_tensor(m[sl], dtype=torch.float64).to(self.device).contiguous()
        d_q = torch.as_tensor(q[sl], dtype=torch.float64).to(self.device).contiguous()
        d_v = torch.as_tensor(vel[sl], dtype=torch.float64).to(self.device).contiguous()
        self._pack(pl.count, d_m, d_q, d_v, is_velocity)
        self._local = (d_m, d_v) if not is_velocity else (d_m, d_v * d_m[:, None])  # m, p
        torch.cuda.current_stream(self.device).synchronize()

    def load_local_device(self, d_m, d_q, d_vel, is_velocity=False):
        """load_local for a caller whose bodies already live on this device (an integrator that
        moved its slice since the last sweep): d_m [count], d_q, d_vel [count][3] device tensors
        of THIS RANK'S slice.  Enqueues the pack on the current stream - no copy, no
        synchronisation - into the row buffer no peer is reading."""
        pl = self.plan
        self.cur = next_row_buffer(self.cur, self._stepped_since_load)
        self._stepped_since_load = False
        if pl.count == 0:
            return
        self._pack(pl.count, d_m, d_q, d_vel, is_velocity)

    def _pack(self, count, d_m, d_q, d_vel, is_velocity):
        """This rank's rows into its own row buffer `cur` (the peers read them from there)."""
        self.ctx.dev_pack(count, d_m.data_ptr(), d_q.data_ptr(), d_vel.data_ptr(), is_velocity,
                          self.regions.packed(self.cur)[self.plan.rank], self._stream())

    def _publish(self):
        """New epoch: tell the peers this rank's rows are in place; returns the shard table."""
        pl, rg = self.plan, self.regions
        self.epoch += 1
        self._stepped_since_load = True
        if pl.world > 1:
            self.ctx.dev_publish(pl.rank, rg.base, self.epoch, self._stream())  # flags at offset 0
        return self.ctx.make_shards(pl.rank, pl.shard, rg.packed(self.cur),
                                    rg.ready_local if pl.world > 1 else 0, self.epoch)

    def _all_reduce(self, t, op):
        if self.plan.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=getattr(dist.ReduceOp, op), group=self.group)
        return t

    def close(self):
        self.torch.cuda.synchronize(self.device)
        barrier = None
        if self.plan.world > 1:
            import torch.distributed as dist
            barrier = lambda: dist.barrier(group=self.group)
            barrier()  # every rank's last sweep has finished reading
        self.regions.close(barrier)


class PeerShardedAccJerk(_PeerSharded):
    """ShardedAccJerk without the all-gather: step() = publish this rank's rows + ONE sweep
    kernel that reads remote source tiles over NVLink (nbk_dev_grad_v_dgrad_v_sum_peer)."""

    def __init__(self, plan: ShardPlan, device, ctx, group=None, exchange=None):
        super().__init__(plan, device, ctx, group, exchange)
        torch = self.torch
        self.g = torch.zeros(max(plan.count, 1), 3, dtype=torch.float64, device=device)
        self.dg = torch.zeros_like(self.g)

    def step(self, eps2: float = 0.0):
        pl = self.plan
        sh = self._publish()
        if pl.count:
            self.ctx.dev_grad_v_dgrad_v_sum_peer(sh, pl.n, eps2, (pl.t0, pl.t1), (0, pl.n),
                                                 self.g.data_ptr(), self.dg.data_ptr(), False,
                                                 self._stream())
        return self.g, self.dg


SYM_BLOCK = 1024  # bodies per block of the pair-once sweep: shards are whole blocks


def sym_item_range(items: int, world: int, rank: int):
    """Rank's contiguous range of the pair-once sweep's work items (block pairs)."""
    return items * rank // world, items * (rank + 1) // world


def sym_item_blocks(n: int, item0: int, item1: int):
    """(I, J) block pairs of items [item0, item1): the upper triangle (diagonal included) of
    ceil(n / 1024) blocks, row-major - the numbering of include/nbk.h."""
    nb = (n + SYM_BLOCK - 1) // SYM_BLOCK
    k = 0
    for i in range(nb):
        for j in range(i, nb):
            if item0 <= k < item1:
                yield i, j
            k += 1


class PeerShardedAccJerkSym(_PeerSharded):
    """The sharded acc+jerk sweep with every unordered pair evaluated ONCE on the whole box
    (hermite.ml:157-166 does the same on one core): the block pairs of the upper triangle are
    dealt out to the ranks in contiguous ranges.  Exchange steps: the packing kernel stores a
    rank's rows into every rank's gathered array over NVLink (the all-gather, fused into the
    pack; ready flags as in the other peer drivers), each rank's kernel runs its block pairs on
    local memory behind the flags (nbk_dev_grad_v_dgrad_v_sym_peer) and leaves partial sums for
    ALL bodies in its region, publishes a second flag, and every rank adds the ranks' sums of
    ITS slice in rank order over peer reads (nbk_dev_sym_gather_peer) - the reduce-scatter,
    deterministic and without a library collective.  reduce="nccl": one NCCL reduce-scatter
    instead.  The plan needs make_plan(..., align=1024).  `sweep`, `reduce_scatter` and `split`
    are injectable for the CPU plumbing tests (an injected reduce_scatter selects that path)."""

    def __init__(self, plan: ShardPlan, device, ctx, group=None, exchange=None,
                 reduce_scatter=None, sweep=None, split=None, reduce="peer"):
        if plan.shard % SYM_BLOCK:
            raise ValueError("the pair-once sweep needs make_plan(..., align=1024)")
        npad = max(plan.npad, 1)
        self.reduce = "nccl" if reduce_scatter is not None else reduce
        # every rank's row buffers hold ALL rows: a rank's block pairs touch rows of every owner
        # several times per sweep, so the rows are gathered once - by the packing kernel's own
        # remote stores (nbk_dev_pack_push) - and the sweep reads local memory; the rank's sums
        # [npad][6] live in the region too, where the peers' gather kernels read them
        super().__init__(plan, device, ctx, group, exchange, rows_per_buffer=npad,
                         extra_bytes=npad * 6 * 8)
        torch = self.torch
        self.sum6 = (torch.zeros(npad, 6, dtype=torch.float64, device=device)
                     if self.reduce == "nccl" or plan.world == 1 else None)
        self.out6 = torch.zeros(max(plan.shard, 1), 6, dtype=torch.float64, device=device)
        self.g = torch.zeros(max(plan.count, 1), 3, dtype=torch.float64, device=device)
        self.dg = torch.zeros_like(self.g)
        nb = (plan.n + SYM_BLOCK - 1) // SYM_BLOCK
        self.items = nb * (nb + 1) // 2
        self.item_range = sym_item_range(self.items, plan.world, plan.rank)
        if reduce_scatter is None:
            def reduce_scatter(out, inp):
                import torch.distributed as dist
                dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.SUM, group=group)
        if sweep is None:
            def sweep(sh, n, eps2, item_range, sum6):
                ptr = sum6 if isinstance(sum6, int) else sum6.data_ptr()
                self.ctx.dev_grad_v_dgrad_v_sym_peer(sh, n, eps2, item_range, ptr, self._stream())
        if split is None:
            def split(src, count, g, dg):
                self.ctx.dev_split6(src.data_ptr(), count, g.data_ptr(), dg.data_ptr(),
                                    self._stream())
        self._reduce_scatter, self._sweep, self._split = reduce_scatter, sweep, split

    def _pack(self, count, d_m, d_q, d_vel, is_velocity):
        """This rank's rows into its place in EVERY rank's gathered array (buffer `cur`)."""
        pl = self.plan
        off = pl.rank * pl.shard * PACKED * 8
        dst = [b + off for b in self.regions.packed(self.cur)]
        self.ctx.dev_pack_push(count, d_m.data_ptr(), d_q.data_ptr(), d_vel.data_ptr(), is_velocity,
                               dst, self._stream())

    def local_shards(self, with_flags=True):
        """Shard table of the LOCAL gathered array: owner r's rows start at r * shard rows."""
        pl, rg = self.plan, self.regions
        base = rg.packed(self.cur)[pl.rank]
        rows = [base + r * pl.shard * PACKED * 8 for r in range(pl.world)]
        flags = rg.ready_local if (with_flags and pl.world > 1) else 0
        return self.ctx.make_shards(pl.rank, pl.shard, rows, flags, self.epoch)

    def sums_ptr(self):
        """Where this rank's sweep leaves its sums [npad][6]: the region (peer reduce) or a
        private tensor (NCCL reduce-scatter, single rank)."""
        if self.sum6 is not None:
            return self.sum6.data_ptr()
        return self.regions.base[self.plan.rank] + self.regions.extra_off

    def step(self, eps2: float = 0.0):
        pl, rg = self.plan, self.regions
        self._publish()
        sh = self.local_shards()
        if self.sum6 is not None:  # NCCL reduce-scatter (or one rank): sums in a private tensor
            self._sweep(sh, pl.n, eps2, self.item_range, self.sum6)
            if pl.world > 1:
                self._reduce_scatter(self.out6, self.sum6)
                src = self.out6
            else:
                src = self.sum6
            if pl.count:
                self._split(src, pl.count, self.g, self.dg)
            return self.g, self.dg
        # peer reduce: sums into the region, second flag, every rank gathers its slice
        self._sweep(sh, pl.n, eps2, self.item_range, self.sums_ptr())
        flags2 = [b + PEER_FLAGS2_OFFSET for b in rg.base]
        self.ctx.dev_publish(pl.rank, flags2, self.epoch, self._stream())
        if pl.count:
            self.ctx.dev_sym_gather_peer(pl.rank, [b + rg.extra_off for b in rg.base],
                                         flags2[pl.rank], self.epoch, (pl.t0, pl.t1),
                                         self.g.data_ptr(), self.dg.data_ptr(), self._stream())
        return self.g, self.dg


class PeerShardedEnergy(_PeerSharded):
    """Energy.energy (energy.ml:57-78) of the sharded system: every rank sweeps
    the potential of its targets over all sources (peer reads), sums its phi and its kinetic
    energies in a fixed order, and ONE all-reduce of two doubles gives (ke, pe) on every rank.
    phi stays on the device for per-body diagnostics (analysis.ml:904-919)."""

    def __init__(self, plan: ShardPlan, device, ctx, group=None, exchange=None):
        super().__init__(plan, device, ctx, group, exchange)
        self.phi = self.torch.zeros(max(plan.count, 1), dtype=self.torch.float64, device=device)

    def step(self, eps2: float = 0.0):
        torch, pl = self.torch, self.plan
        sh = self._publish()
        part = torch.zeros(2, dtype=torch.float64, device=self.device)
        if pl.count:
            self.ctx.dev_v_sum_peer(sh, pl.n, eps2, (pl.t0, pl.t1), (0, pl.n), self.phi.data_ptr(),
                                    False, self._stream())
            d_m, d_p = self._local
            part[0] = ((d_p * d_p).sum(dim=1) / (2.0 * d_m)).sum()
            part[1] = 0.5 * self.phi[:pl.count].sum()
        self._all_reduce(part, "SUM")
        ke, pe = (float(x) for x in part.tolist())
        return ke, pe


class PeerShardedKick(_PeerSharded):
    """The all-pairs Leapfrog.kick sweep (leapfrog.ml:151-155, pre-kick order) of the sharded
    system: dv and the per-body min time-scale for this rank's targets; the global minimum
    time-scale (the step a shared-step integrator would take) is one all-reduce(min)."""

    def __init__(self, plan: ShardPlan, device, ctx, group=None, exchange=None):
        super().__init__(plan, device, ctx, group, exchange)
        torch = self.torch
        self.dv = torch.zeros(max(plan.count, 1), 3, dtype=torch.float64, device=device)
        self.tmin = torch.full((max(plan.count, 1),), float("inf"), dtype=torch.float64, device=device)

    def load_local(self, m, q, v):
        super().load_local(m, q, v, is_velocity=True)

    def step(self, eta: float, dt: float):
        torch, pl = self.torch, self.plan
        sh = self._publish()
        if pl.count:
            self.ctx.dev_kick_sum_peer(sh, pl.n, eta, dt, (pl.t0, pl.t1), (0, pl.n),
                                       self.dv.data_ptr(), self.tmin.data_ptr(), self._stream())
        gmin = torch.full((1,), float("inf"), dtype=torch.float64, device=self.device)
        if pl.count:
            ok = self.tmin[:pl.count]
            ok = ok[~torch.isnan(ok)]
            if ok.numel():
                gmin[0] = ok.min()
        self._all_reduce(gmin, "MIN")
        return self.dv, self.tmin, float(gmin.item())


class PeerShardedTangent:
    """ShardedTangent over the peer-memory exchange: bodies and the K direction arrays are read
    from their owners by the tangent sweep itself."""

    def __init__(self, plan: ShardPlan, K: int, device, ctx, group=None, exchange=None):
        import torch
        self.torch, self.plan, self.K, self.device, self.ctx = torch, plan, K, device, ctx
        self.group = group
        self.regions = PeerRegions(plan, ctx, K, group, exchange)
        self.gt = torch.zeros(K, max(plan.count, 1), 3, dtype=torch.float64, device=device)
        self.dgt = torch.zeros_like(self.gt)
        self.epoch, self.cur = 0, 0
        self._stepped_since_load = True

    def load_local(self, m, q, p, dq, dp):
        torch, pl, rg = self.torch, self.plan, self.regions
        self.cur = next_row_buffer(self.cur, self._stepped_since_load)
        self._stepped_since_load = False
        if pl.count == 0:
            return
        sl = slice(pl.t0, pl.t1)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        d_m = torch.as_tensor(m[sl]).to(self.device).contiguous()
        d_q = torch.as_tensor(q[sl]).to(self.device).contiguous()
        d_p = torch.as_tensor(p[sl]).to(self.device).contiguous()
        self.ctx.dev_pack(pl.count, d_m.data_ptr(), d_q.data_ptr(), d_p.data_ptr(), False,
                          rg.packed(self.cur)[pl.rank], stream)
        keep = [d_m, d_q, d_p]
        dir_bytes = pl.shard * PACKED * 8
        for k in range(self.K):
            d_dq = torch.as_tensor(dq[k][sl]).to(self.device).contiguous()
            d_dp = torch.as_tensor(dp[k][sl]).to(self.device).contiguous()
            self.ctx.dev_pack_aux(pl.count, d_m.data_ptr(), d_dq.data_ptr(), d_dp.data_ptr(),
                                  rg.aux(self.cur)[pl.rank] + k * dir_bytes, stream)
            keep += [d_dq, d_dp]
        torch.cuda.current_stream(self.device).synchronize()

    def load_local_predicted(self, t, m, q, p, f, df, t_tan, q_tan, p_tan, f_tan, df_tan, nt,
                             nt_tan=None):
        """The dual-number Hermite step's exchange (DFloat_hermite.advance_body,
        dFloat_hermite.ml:115-135, all bodies as targets): this rank PREDICTS its own slice to
        the dual time nt - values and the K tangents of (t, q, p, f, df) - straight into its
        region (nbk_dev_predict_pack / _tangent); step() then sweeps as usual.  Host arrays of
        the FULL system: t, m [n]; q, p, f, df [n][3]; t_tan [K][n] or None; *_tan [K][n][3];
        nt_tan [K] or None."""
        torch, pl, rg = self.torch, self.plan, self.regions
        self.cur = next_row_buffer(self.cur, self._stepped_since_load)
        self._stepped_since_load = False
        if pl.count == 0:
            return
        sl = slice(pl.t0, pl.t1)
        stream = torch.cuda.current_stream(self.device).cuda_stream

        def up(a):
            return torch.as_tensor(a, dtype=torch.float64).to(self.device).contiguous()
        d = [up(a[sl]) for a in (t, m, q, p, f, df)]
        self.ctx.dev_predict_pack(pl.count, *[x.data_ptr() for x in d], nt,
                                  rg.packed(self.cur)[pl.rank], stream)
        d_tt = up(t_tan[:, sl]) if t_tan is not None else None
        d_tan = [up(a[:, sl]) for a in (q_tan, p_tan, f_tan, df_tan)]
        d_ntt = up(nt_tan) if nt_tan is not None else None
        self.ctx.dev_predict_pack_tangent(
            pl.count, self.K, *[x.data_ptr() for x in d], nt,
            d_tt.data_ptr() if d_tt is not None else 0, *[x.data_ptr() for x in d_tan],
            d_ntt.data_ptr() if d_ntt is not None else 0, rg.aux(self.cur)[pl.rank],
            pl.shard * PACKED, stream)
        torch.cuda.current_stream(self.device).synchronize()

    def step(self, eps2: float = 0.0):
        pl, rg = self.plan, self.regions
        self.epoch += 1
        self._stepped_since_load = True
        st = self.torch.cuda.current_stream(self.device).cuda_stream
        if pl.world > 1:
            self.ctx.dev_publish(pl.rank, rg.base, self.epoch, st)
        if pl.count:
            sh = self.ctx.make_shards(pl.rank, pl.shard, rg.packed(self.cur),
                                      rg.ready_local if pl.world > 1 else 0, self.epoch,
                                      aux=rg.aux(self.cur))
            self.ctx.dev_grad_v_dgrad_v_sum_tangent_peer(
                sh, pl.n, self.K, eps2, (pl.t0, pl.t1), (0, pl.n), self.gt.data_ptr(),
                self.dgt.data_ptr(), st)
        return self.gt, self.dgt

    def close(self):
        self.torch.cuda.synchronize(self.device)
        barrier = None
        if self.plan.world > 1:
            import torch.distributed as dist
            barrier = lambda: dist.barrier(group=self.group)
            barrier()
        self.regions.close(barrier)
