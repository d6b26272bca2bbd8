"""Target-sharded, source-replicated step on several GPUs: the host side.

The data plane lives in libvvgpu (vvgpu_shard.cuh, vvgpu_comm.h): every rank holds the full particle set and rebuilds
the (deterministic) tree, the leaf groups are dealt block-cyclically over the ranks, each rank computes epsilon,
convective and diffusive velocities for the particles of ITS groups, and the library gathers the results on its own
stream inside the calls (`_1_eps` and the merge columns after epsilon, `v` after the velocity phases, `fric` summed).
What is left for the host layer is plumbing:

  * one process per GPU (torchrun): carry the NCCL unique id from rank 0 to the other ranks over torch.distributed
    (any backend: the id is 128 bytes of host memory) and hand it to `vvgpu_comm_init`;
  * one process, several ranks (`capi.group_create`, devices may repeat): drive every context from its own thread,
    all making the same calls.
"""
import threading

import numpy as np


def init_comm(ctx, rank, world, group=None):
    """vvgpu_comm_init on every rank of a torch.distributed job"""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    from . import capi
    ident = capi.comm_unique_id() if rank == 0 else bytes(128)
    box = [ident]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init(rank, world, box[0])


def owned_groups(ngroups, rank, world):
    """leaf groups rank `rank` computes (the library's own rule, vvgpu_shard_owner)"""
    from . import capi
    return [g for g in range(ngroups) if capi.shard_owner(g, world) == rank]


class ShardedStep:
    """One hot-path step (vvflow.cpp:246-257); the same calls on every rank, the exchanges happen inside them."""

    def __init__(self, ctx, rank=0, world=1, device=None, group=None):
        self.ctx, self.rank, self.world = ctx, rank, world
        if world > 1 and ctx.comm_info()[1] == 1:
            init_comm(ctx, rank, world, group)

    def step(self, far, min_node, max_node, merge, inf_vx, inf_vy, dt, re, viscous=True, want_fric=False):
        ctx = self.ctx
        ctx.tree_build(far, min_node, max_node)
        merged = ctx.epsilon(merge)
        ctx.convective(inf_vx, inf_vy, dt)
        fric = ctx.diffusive(re, want_fric=want_fric) if viscous else None
        ctx.tree_destroy()
        out = ctx.move_and_clean(dt)
        out["merged"] = merged
        out["fric"] = fric
        return out


def run_group(ctxs, fn):
    """call fn(rank, ctx) on one thread per context of an in-process group; returns the results in rank order and
    re-raises the first exception"""
    res, err = [None] * len(ctxs), [None] * len(ctxs)

    def work(r):
        try:
            res[r] = fn(r, ctxs[r])
        except BaseException as e:   # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(len(ctxs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return res
