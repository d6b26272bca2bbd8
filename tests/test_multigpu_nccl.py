"""Two-GPU NCCL test of the sharded step (skipped with fewer than two CUDA devices): every target's epsilon and
velocity are computed by exactly one rank from the same per-group lists as on a single GPU, so after the exchange
the step must be BIT-IDENTICAL to the single-GPU step, on every rank. Run with `gpurun --gpus 2`."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, xyg, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from vvflow_b200 import capi, multigpu
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    ctx = capi.Context(rank)
    st = multigpu.ShardedStep(ctx, rank, world, f"cuda:{rank}")
    ctx.set_particles_xyg(xyg)
    merged = []
    for _ in range(steps):
        out = st.step(8, 0.0, float(np.finfo(np.float64).max), True, 1.0, 0.0, 0.05, 600.0)
        merged.append(out["merged"])
    ctx.synchronize()
    torch.cuda.synchronize()
    q.put((rank, ctx.get_particles(), merged, st.bounds.tolist()))
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


def _single(xyg, steps):
    sys.path.insert(0, ROOT)
    from vvflow_b200 import capi, multigpu
    ctx = capi.Context(0)
    st = multigpu.ShardedStep(ctx, 0, 1, "cuda:0")
    ctx.set_particles_xyg(xyg)
    merged = [st.step(8, 0.0, float(np.finfo(np.float64).max), True, 1.0, 0.0, 0.05, 600.0)["merged"] for _ in range(steps)]
    out = ctx.get_particles()
    ctx.close()
    return out, merged


@pytest.mark.parametrize("sign", ["same", "mixed"])
def test_two_gpus_match_one(sign):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    xyg = cases.cloud(60000, "gauss", sign, seed=61)   # "mixed": merges happen, the replicated replay path runs
    steps = 2
    want, wmerged = _single(xyg, steps)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200) + (0 if sign == "same" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, xyg, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
    for rank, got, merged, bounds in res:
        assert merged == wmerged, (rank, merged, wmerged)
        assert got.shape == want.shape, (rank, got.shape, want.shape)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), f"rank {rank} differs from the single-GPU step"
        assert bounds[0][0] == 0 and bounds[0][1] == bounds[1][0]
    if sign == "mixed":
        assert sum(wmerged) > 0
