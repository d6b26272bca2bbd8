"""world_size-2 gloo test (CPU) of the multi-GPU host logic in vvflow_b200/multigpu.py: slice
bookkeeping, tiling check, uneven all-gather, and the probe -> replay decision for merging. The CUDA
context is replaced by a numpy double with the same method names; the exchange code is the real one."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeCtx:
    """stands in for capi.Context: particle i 'computes' f(i) for the targets of its slice only"""

    def __init__(self, n, rank, world, candidates_on=None):
        self.nn, self.rank, self.world = n, rank, world
        self.arr = [np.zeros(n) for _ in range(6)]
        self.candidates_on = candidates_on
        self.calls = []
        cuts = [0, n // 3, n] if world == 2 else np.linspace(0, n, world + 1).astype(int).tolist()
        self.first, self.last = cuts[rank], cuts[rank + 1]

    n = property(lambda self: self.nn)

    def set_shard(self, rank, world): pass
    def tree_build(self, *a): self.calls.append("build")
    def tree_destroy(self): self.calls.append("destroy")
    def shard_range(self): return self.first, self.last
    def synchronize(self): pass
    def tensors(self): return [torch.from_numpy(a) for a in self.arr]

    def epsilon_probe(self):
        self.calls.append("probe")
        self.arr[5][self.first:self.last] = 1.0 + np.arange(self.first, self.last)
        return 3 if self.candidates_on == self.rank else 0

    def epsilon(self, merge):
        self.calls.append("eps_replicated" if merge else "eps")
        if merge:  # replicated replay: every rank computes everything
            self.arr[5][:] = 100.0 + np.arange(self.nn)
            return 7
        self.arr[5][self.first:self.last] = 1.0 + np.arange(self.first, self.last)
        return 0

    def convective(self, *a):
        self.arr[3][self.first:self.last] = 2.0 * np.arange(self.first, self.last)
        self.arr[4][self.first:self.last] = -1.0 * np.arange(self.first, self.last)

    def diffusive(self, re, want_fric=False):
        self.arr[3][self.first:self.last] += 0.5

    def move_and_clean(self, dt):
        return {"cleaned": 0}


def _worker(rank, world, port, n, candidates_on, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vvflow_b200 import multigpu
    ctx = FakeCtx(n, rank, world, candidates_on)
    step = multigpu.ShardedStep(ctx, rank, world, "cpu")
    out = step.step(8, 0.0, 1e300, True, 1.0, 0.0, 0.005, 1000.0)
    ok = True
    i = np.arange(n)
    if candidates_on is None:
        ok &= bool(np.array_equal(ctx.arr[5], 1.0 + i)) and "eps_replicated" not in ctx.calls
    else:
        ok &= bool(np.array_equal(ctx.arr[5], 100.0 + i)) and "eps_replicated" in ctx.calls and out["merged"] == 7
    ok &= bool(np.array_equal(ctx.arr[3], 2.0 * i + 0.5)) and bool(np.array_equal(ctx.arr[4], -1.0 * i))
    # a broken tiling must be detected
    try:
        multigpu.check_tiling(np.array([[0, 5], [6, n]]), n)
        ok = False
    except RuntimeError:
        pass
    q.put((rank, ok, step.bounds.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("candidates_on", [None, 1])
def test_sharded_step_two_ranks(candidates_on):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (0 if candidates_on is None else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1000, candidates_on, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2] == [[0, 333], [333, 1000]]
