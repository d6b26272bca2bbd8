"""world_size-2 gloo test (CPU) of the multi-GPU HOST logic that is left in vvflow_b200/multigpu.py now that the data
plane lives in the library: the NCCL unique id travels from rank 0 to every rank, every rank hands the SAME id and its
own (rank, world) to vvgpu_comm_init, and the ranks then make identical call sequences. The CUDA context is replaced by
a recording double; the ownership rule is the library's own (vvgpu_shard_owner needs no device)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeCtx:
    def __init__(self):
        self.calls, self.comm = [], None

    def comm_info(self):
        return (0, 1, 0) if self.comm is None else (self.comm[0], self.comm[1], 1)

    def comm_init(self, rank, nranks, ident):
        self.comm = (rank, nranks, bytes(ident))

    def tree_build(self, *a): self.calls.append(("build",) + a)
    def epsilon(self, merge): self.calls.append(("eps", merge)); return 5
    def convective(self, *a): self.calls.append(("conv",) + a)
    def diffusive(self, re, want_fric=False): self.calls.append(("diff", re)); return None
    def tree_destroy(self): self.calls.append(("destroy",))
    def move_and_clean(self, dt): self.calls.append(("move", dt)); return {"cleaned": 0}


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    from vvflow_b200 import capi, multigpu
    dist.init_process_group("gloo", rank=rank, world_size=world)
    capi.comm_unique_id = lambda: bytes(range(128))       # no NCCL needed on the CPU box: the id is opaque bytes
    ctx = FakeCtx()
    st = multigpu.ShardedStep(ctx, rank, world)
    out = st.step(8, 0.0, 1e300, True, 1.0, 0.0, 0.05, 600.0)
    q.put((rank, ctx.comm, ctx.calls, out["merged"]))
    dist.barrier()
    dist.destroy_process_group()


def test_unique_id_reaches_every_rank_and_calls_are_identical():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert [r[1][:2] for r in res] == [(0, 2), (1, 2)]
    assert res[0][1][2] == res[1][1][2] == bytes(range(128))
    assert res[0][2] == res[1][2] and [c[0] for c in res[0][2]] == ["build", "eps", "conv", "diff", "destroy", "move"]
    assert res[0][3] == res[1][3] == 5


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_ownership_tiles_the_groups(world):
    """every leaf group has exactly one owner, pieces of consecutive groups go round the ranks"""
    sys.path.insert(0, ROOT)
    from vvflow_b200 import capi, multigpu
    blk = capi.shard_block()
    ng = 1003
    owned = [multigpu.owned_groups(ng, r, world) for r in range(world)]
    allg = sorted(g for o in owned for g in o)
    assert allg == list(range(ng))
    sizes = [len(o) for o in owned]
    assert max(sizes) - min(sizes) <= blk
    for r, o in enumerate(owned):
        assert all((g // blk) % world == r for g in o)
