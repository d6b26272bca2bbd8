"""Two-GPU NCCL test of the sharded step (skipped with fewer than two CUDA devices; run with `gpurun --gpus 2`): one
process per GPU, the library's own ncclAllGather exchanges. Every target's epsilon and velocity are computed by exactly
one rank from the same per-group lists as on a single GPU, so after the exchanges the state must be BIT-IDENTICAL to
the single-GPU step, on every rank; fric is a sum over ranks (1e-10)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _setup(ctx, xyg, with_body):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    from vvflow_b200 import vvhd
    bodies = [cases.cylinder(0.5, 350)] if with_body else []
    S = vvhd.Space(ctx=ctx)
    S.BodyList = bodies
    ctx.set_bodies(*S._pack_bodies())
    ctx.set_particles_xyg(xyg)
    return cases.tree_params(bodies)


def _steps(ctx, tree, steps):
    log = []
    for _ in range(steps):
        ctx.tree_build(8, tree[0], tree[1])
        merged = ctx.epsilon(True)
        ctx.convective(1.0, 0.0, 0.05)
        fric = ctx.diffusive(600.0, want_fric=True)
        ctx.tree_destroy()
        out = ctx.move_and_clean(0.05)
        log.append((merged, fric, out["cleaned"]))
    return log, ctx.get_particles()


def _worker(rank, world, port, xyg, with_body, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from vvflow_b200 import capi, multigpu
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # plumbing only: carries the NCCL id
    ctx = capi.Context(rank)
    multigpu.init_comm(ctx, rank, world)
    tree = _setup(ctx, xyg, with_body)
    log, got = _steps(ctx, tree, steps)
    q.put((rank, got, log))
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


@pytest.mark.parametrize("case", ["same", "mixed", "cylinder"])
def test_two_gpus_match_one(case):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    from vvflow_b200 import capi
    with_body = case == "cylinder"
    xyg = cases.around_cylinder(40000, sign="mixed", seed=62) if with_body else cases.cloud(60000, "gauss", case, seed=61)
    steps = 2
    one = capi.Context(0)
    want_log, want = _steps(one, _setup(one, xyg, with_body), steps)
    one.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200) + {"same": 0, "mixed": 1, "cylinder": 2}[case]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, xyg, with_body, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
    for rank, got, log in res:
        assert [l[0] for l in log] == [l[0] for l in want_log] and [l[2] for l in log] == [l[2] for l in want_log]
        assert got.shape == want.shape, (rank, got.shape, want.shape)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), f"rank {rank} differs from the single-GPU step"
        if with_body:
            for a, b in zip(log, want_log):
                cases.check_close(a[1], b[1], 1e-10, "fric over two ranks")
    if case != "same":
        assert sum(l[0] for l in want_log) > 0
