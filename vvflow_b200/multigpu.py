"""Target-sharded, source-replicated step over torch.distributed (one process per GPU).

Every rank holds the full particle set and rebuilds the (deterministic) tree itself; the leaf groups
are cut into contiguous slices balanced by work units, and each rank computes epsilon, convective and
diffusive velocities only for the particles of its slice. Between phases the slices are all-gathered
(NCCL over NVLink on the GPU box, gloo in the CPU tests), so that the next phase again sees every
source: `_1_eps` (8 B/particle) after epsilon, `v` (16 B/particle) after the velocity phases. The
order-dependent merge replay and move_and_clean are replicated (they are cheap and deterministic),
so all ranks stay bit-identical without exchanging positions. SURVEY.md §8(e).
"""
import numpy as np
import torch
import torch.distributed as dist


def slice_bounds(first, last, group=None):
    """all ranks' [first, last) particle ranges, as a (world, 2) int64 array"""
    world = dist.get_world_size(group)
    mine = torch.tensor([first, last], dtype=torch.int64)
    backend = dist.get_backend(group)
    if backend == "nccl":
        mine = mine.cuda()
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return torch.stack(out).cpu().numpy()


def check_tiling(bounds, n):
    """slices must tile [0, n) in rank order (they are cut from the same deterministic group list)"""
    b = np.asarray(bounds)
    if b[0, 0] != 0 or b[-1, 1] != n or np.any(b[1:, 0] != b[:-1, 1]) or np.any(b[:, 1] < b[:, 0]):
        raise RuntimeError(f"shard slices do not tile [0,{n}): {b.tolist()}")


def allgather_padded(fulls, bounds, rank, group=None):
    """Every tensor in `fulls` is a length-n tensor of which this rank owns [bounds[rank,0], bounds[rank,1]); on
    return every rank holds every slice of every tensor. ONE collective for all of them: the (uneven) slices are
    padded to the longest one, gathered with all_gather (equal sizes), and written back with one concatenation
    per tensor. One broadcast per rank and array cost world x len(fulls) collectives per phase boundary, which
    dominated the step at 8 GPUs."""
    world = bounds.shape[0]
    lens = [int(bounds[r, 1] - bounds[r, 0]) for r in range(world)]
    L = max(max(lens), 1)
    k = len(fulls)
    a, b = int(bounds[rank, 0]), int(bounds[rank, 1])
    send = torch.zeros((k, L), dtype=fulls[0].dtype, device=fulls[0].device)
    for i, f in enumerate(fulls):
        send[i, : b - a] = f[a:b]
    if dist.get_backend(group) == "nccl":
        recv = torch.empty((world, k, L), dtype=send.dtype, device=send.device)
        dist.all_gather_into_tensor(recv, send, group=group)
        parts = [recv[r] for r in range(world)]
    else:
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send, group=group)
    for i, f in enumerate(fulls):
        torch.cat([parts[r][i, : lens[r]] for r in range(world)], out=f)
    return fulls


class DevArray:
    """torch view of a raw device pointer owned by libvvgpu (no copy)"""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def dev_tensor(ptr, n, device):
    if n == 0:
        return torch.empty(0, dtype=torch.float64, device=device)
    return torch.as_tensor(DevArray(ptr, n), device=device)


class ShardedStep:
    """One hot-path step (vvflow.cpp:246-257) on `world` GPUs."""

    def __init__(self, ctx, rank, world, device, group=None):
        self.ctx, self.rank, self.world, self.device, self.group = ctx, rank, world, device, group
        ctx.set_shard(rank, world)
        # On a GPU the collectives are enqueued relative to the LIBRARY's stream (torch sees it as an external
        # stream): NCCL orders itself against that stream with events, so a phase boundary needs no host
        # synchronisation at all.
        self.ext = None
        if world > 1 and torch.device(device).type == "cuda" and hasattr(ctx, "stream"):
            self.ext = torch.cuda.ExternalStream(ctx.stream(), device=torch.device(device))

    def _tensors(self):
        if hasattr(self.ctx, "tensors"):       # test doubles hand out CPU tensors directly
            return self.ctx.tensors()
        ptrs, n = self.ctx.arrays_dev()
        return [dev_tensor(p, n, self.device) for p in ptrs]

    def _gather(self, which):
        if self.ext is not None:
            with torch.cuda.stream(self.ext):
                ts = self._tensors()
                allgather_padded([ts[k] for k in which], self.bounds, self.rank, self.group)
            return
        self.ctx.synchronize()                 # CPU test doubles / no stream handle: plain host ordering
        ts = self._tensors()
        allgather_padded([ts[k] for k in which], self.bounds, self.rank, self.group)
        if torch.device(self.device).type == "cuda":
            torch.cuda.synchronize(self.device)

    def _bounds(self):
        if hasattr(self.ctx, "shard_bounds"):  # the tree and the cut are replicated: no communication needed
            return self.ctx.shard_bounds(self.world)
        first, last = self.ctx.shard_range()
        return slice_bounds(first, last, self.group)

    def step(self, far, min_node, max_node, merge, inf_vx, inf_vy, dt, re, viscous=True):
        ctx = self.ctx
        ctx.tree_build(far, min_node, max_node)
        if self.world > 1:
            self.bounds = self._bounds()
            check_tiling(self.bounds, ctx.n)
            merged = 0
            if merge:
                # merging is order-dependent: probe the slice; only if some rank has a candidate
                # does every rank replay the merges (replicated, bit-identical everywhere)
                cand = torch.tensor([ctx.epsilon_probe()], dtype=torch.int64, device=self.device)
                dist.all_reduce(cand, group=self.group)
                if int(cand.item()) > 0:
                    merged = ctx.epsilon(True)
                else:
                    self._gather([5])          # _1_eps of the other ranks' targets
            else:
                ctx.epsilon(False)
                self._gather([5])
        else:
            merged = ctx.epsilon(merge)
        ctx.convective(inf_vx, inf_vy, dt)
        if viscous:
            ctx.diffusive(re, want_fric=False)
        if self.world > 1:
            self._gather([3, 4])               # v
        ctx.tree_destroy()
        out = ctx.move_and_clean(dt)
        out["merged"] = merged
        return out
