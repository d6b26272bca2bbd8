"""The sharded step with several ranks in ONE process (capi.group_create, one thread per rank, peer-to-peer copies
between the ranks): 2, 3 and 4 ranks on cuda:0, so the whole multi-rank code path (block-cyclic ownership, per-phase
gathers, the merge fixed point iterated on gathered columns, fric summed over ranks) is exercised on a one-GPU box.
Every target is computed by exactly one rank from the same per-group lists as on a single GPU, so the state must be
BIT-IDENTICAL to the single-context step on every rank; fric is a sum over ranks (1e-10)."""
import numpy as np
import pytest

import cases
from cases import check_close

pytestmark = pytest.mark.gpu
DBL_MAX = float(np.finfo(np.float64).max)


def _space(ctx, xyg, bodies):
    from vvflow_b200 import vvhd
    S = vvhd.Space(ctx=ctx)
    S.BodyList = bodies
    segs, bd = S._pack_bodies()
    ctx.set_bodies(segs, bd)
    ctx.set_particles_xyg(xyg)
    return S


def _steps(ctx, nsteps, tree, sinks=None, re=600.0, dt=0.05, snapshots=None):
    """nsteps hot-path steps; returns per step (merged, fric, cleaned, gsum) and the final records"""
    log = []
    for k in range(nsteps):
        ctx.tree_build(8, tree[0], tree[1])
        merged = ctx.epsilon(True)
        ctx.convective(1.0, 0.25, dt, sinks)
        fric = ctx.diffusive(re, want_fric=True)
        if snapshots is not None:
            snapshots.append(ctx.get_particles())      # velocities before the move: forces the v gather
        ctx.tree_destroy()
        out = ctx.move_and_clean(dt)
        log.append((merged, None if fric is None else fric.copy(), out["cleaned"], out["gsum"].copy(), out["fdt_dead"].copy()))
    return log, ctx.get_particles()


def _run_case(xyg, bodies, nranks, nsteps=2, sinks=None):
    from vvflow_b200 import capi, multigpu
    tree = cases.tree_params(bodies)
    one = capi.Context(0)
    _space(one, xyg, bodies)
    snap1 = []
    want_log, want = _steps(one, nsteps, tree, sinks, snapshots=snap1)
    one.close()
    ctxs = capi.group_create([0] * nranks)
    try:
        for c in ctxs:
            _space(c, xyg, bodies)
        snaps = [[] for _ in ctxs]
        res = multigpu.run_group(ctxs, lambda r, c: _steps(c, nsteps, tree, sinks, snapshots=snaps[r]))
    finally:
        for c in ctxs:
            c.close()
    for r, (log, got) in enumerate(res):
        assert got.shape == want.shape, (r, got.shape, want.shape)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), f"rank {r} of {nranks} differs from the single-context step"
        for k in range(nsteps):
            assert np.array_equal(snaps[r][k].view(np.uint64), snap1[k].view(np.uint64)), f"rank {r}: state before move of step {k}"
            assert log[k][0] == want_log[k][0] and log[k][2] == want_log[k][2], (r, k, log[k][0], want_log[k][0])
            if bodies:
                check_close(log[k][1], want_log[k][1], 1e-10, f"fric, rank {r}")
                # gsum / fdt_dead are floating-point atomics over the removed particles: order-dependent sums
                check_close(log[k][3], want_log[k][3], 1e-10, f"gsum, rank {r}")
                check_close(log[k][4].ravel(), want_log[k][4].ravel(), 1e-10, f"fdt_dead, rank {r}")
    return want_log


@pytest.mark.parametrize("nranks", [2, 3, 4])
@pytest.mark.parametrize("sign", ["same", "mixed"])
def test_group_cloud(nranks, sign):
    log = _run_case(cases.cloud(40000, "gauss", sign, seed=61 + nranks), [], nranks)
    if sign == "mixed":
        assert sum(l[0] for l in log) > 0     # merges happened: the fixed point ran on gathered columns


@pytest.mark.parametrize("nranks", [2, 4])
def test_group_cylinder_with_merges_and_fric(nranks):
    """wall epsilon restriction and merge criterion (per-leaf wall parameters of the rank's own leaves only), segment
    diffusion + fric summed over ranks, in-body removal; sinks through process_all_lists"""
    body = cases.cylinder(0.5, 350)
    sinks = np.array([[2.0, 0.3, 0.4], [-1.5, 0.1, -1.2]])
    log = _run_case(cases.around_cylinder(30000, sign="mixed", seed=71), [body], nranks, sinks=sinks)
    assert sum(l[0] for l in log) > 0 and sum(l[2] for l in log) > 0


def test_group_few_groups():
    """fewer pieces than ranks: some ranks own nothing"""
    _run_case(cases.cloud(300, "gauss", "mixed", seed=5), [], 4, nsteps=1)
    _run_case(np.zeros((0, 3)), [], 2, nsteps=1)
    # ... also with walls (the per-leaf wall passes run over the rank's work units: none), and a body with no vortices
    # at all, which is step 0 of every vvflow run
    body = cases.cylinder(0.5, 350)
    _run_case(cases.around_cylinder(200, sign="mixed", seed=9), [body], 4, nsteps=2)
    _run_case(np.zeros((0, 3)), [body], 2, nsteps=1)


def test_group_slice_upload():
    """vvgpu_set_particles_slice: one slice per rank, gathered over the transport"""
    from vvflow_b200 import capi, multigpu
    n = 10007
    rec = np.random.default_rng(3).standard_normal((n, 6))
    ctxs = capi.group_create([0, 0, 0])
    try:
        def up(r, c):
            lo, hi = n * r // 3, n * (r + 1) // 3
            c.set_particles_slice(rec[lo:hi], lo, n)
            return c.get_particles()
        for got in multigpu.run_group(ctxs, up):
            assert np.array_equal(got, rec)
    finally:
        for c in ctxs:
            c.close()
