"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): tree build, particle order, interaction lists, epsilon and merge
decisions BIT-EXACT; velocities and integral sums within 1e-10 relative, in fp64.
The oracle here is oracle/port (C restatement, itself pinned bit-exact to the compiled reference
by tests/test_oracle_port.py); where oracle/_ref travelled to the box the reference build is
checked directly as well.
"""
import numpy as np
import pytest

import cases
from cases import check_close, relerr, same

pytestmark = pytest.mark.gpu

VTOL = 1e-10  # relative tolerance on velocities / integral sums stated by north_star


def run_pair(ctx, port, xyg, bodies=(), merge=True, re=600.0, dt=0.05, inf=(1.0, 0.0), far=8, stop=None, tree=None,
             sinks=None, lists=True):
    """Drive oracle and GPU through the hot path (vvflow.cpp:246-257), comparing after every phase."""
    bodies = list(bodies)
    mn, mx = tree if tree is not None else cases.tree_params(bodies)
    pb = cases.port_bodies(port, bodies)
    P = port.Port(xyg=xyg, bodies=pb)
    from vvflow_b200 import vvhd
    S = vvhd.Space(ctx=ctx)
    S.VortexList = xyg
    S.BodyList = bodies
    S.re, S.dt, S.inf_vx, S.inf_vy = re, dt, inf[0], inf[1]
    if sinks is not None:
        S.SourceList = np.asarray(sinks, dtype=np.float64)
    tr = vvhd.TSortedTree(S, far, mn, mx)
    eps, conv, diff, flow = vvhd.MEpsilonFast(S, tr), vvhd.MConvectiveFast(S, tr), vvhd.MDiffusiveFast(S, tr), vvhd.MFlowmove(S)
    out = {}
    try:
        # ---- tree
        P.tree_build(far, mn, mx)
        tr.build()
        d1, i1, nl1 = P.tree_export()
        d2, i2, nl2 = ctx.tree_export()
        assert nl1 == nl2 and d1.shape == d2.shape, (nl1, nl2, d1.shape, d2.shape)
        assert same(d2[:, :4], d1[:, :4]), "node boxes differ"
        assert same(i2[:, [0, 1, 2, 3, 4, 5]], i1[:, [0, 1, 6, 7, 8, 9]]), "node ranges / children / leaf order differ"
        assert same(d2[:, 4:], d1[:, 4:]), "centres of mass differ"
        assert same(S.VortexList[:, :3], P.rec48()[:, :3]), "permuted particle order differs"
        assert same(ctx.get_permutation(), P.orig[: P.n]), "permutation differs"
        if lists:   # (at the benchmarked sizes the per-leaf lists are GBs: the pair counts below still compare them)
            l1, l2 = P.tree_lists(), ctx.tree_lists()
            for a, b, name in zip(l2, l1, ("near_ptr", "near_idx", "far_ptr", "far_idx")):
                assert same(a, b), f"interaction lists differ: {name}"
        if bodies:
            s1, s2 = P.tree_leaf_segments(), ctx.tree_leaf_segments()
            assert same(s2[0], s1[0]) and same(s2[1], s1[1]), "leaf segment lists differ"
        np1, nf1 = P.count_interactions()
        np2, nf2 = ctx.count_interactions()
        assert np1 == np2 and nf1 == nf2, ("group lists disagree with per-leaf lists", np1, np2, nf1, nf2)
        out["near_pairs"], out["far_nodes"], out["leaves"] = np2, nf2, nl2
        if stop == "tree":
            return out
        # ---- epsilon (+ merge)
        m1 = P.epsilon(merge)
        eps.CalcEpsilonFast(merge)
        assert eps.Merged() == m1, ("merge count", eps.Merged(), m1)
        a, b = S.VortexList, P.rec48()
        assert same(a[:, :3], b[:, :3]), "merged positions / circulations differ"
        assert same(a[:, 5], b[:, 5]), f"epsilon differs, max rel {relerr(a[:, 5], b[:, 5])}"
        out["merged"] = m1
        if stop == "eps":
            return out
        # ---- convective
        P.convective(inf[0], inf[1], dt, sinks=sinks)
        conv.process_all_lists()
        a, b = S.VortexList, P.rec48()
        out["conv_err"] = check_close(a[:, 3:5], b[:, 3:5], VTOL, "convective velocity")
        # ---- diffusive
        if np.isfinite(re):
            P.diffusive(re)
            diff.process_vort_list()
            a, b = S.VortexList, P.rec48()
            out["diff_err"] = check_close(a[:, 3:5], b[:, 3:5], VTOL, "convective + diffusive velocity")
            if bodies:
                fr = np.concatenate([bd.fric for bd in bodies])
                out["fric_err"] = check_close(fr, pb.a["fric"], VTOL, "fric")
        # ---- move and clean
        P.tree_destroy()
        tr.destroy()
        n1, c1 = P.move_and_clean(dt)
        c2 = flow.move_and_clean(True)
        a, b = S.VortexList, P.rec48()
        assert a.shape[0] == n1 and c2 == c1, ("survivors / cleaned", a.shape[0], n1, c2, c1)
        assert same(a[:, 2], b[:, 2]) and same(a[:, 5], b[:, 5]), "survivor g / eps differ"
        assert same(a[:, 3:5], b[:, 3:5]), "v not zeroed"
        out["pos_err"] = check_close(a[:, :2], b[:, :2], VTOL, "advected positions")
        if bodies:
            gs = np.concatenate([bd.gsum for bd in bodies])
            assert relerr(gs, pb.a["gsum"]) <= VTOL
            fd = np.concatenate([bd.fdt_dead for bd in bodies])
            assert relerr(fd, pb.a["fdt_dead"][: fd.shape[0]]) <= VTOL
            gd = np.array([bd.g_dead for bd in bodies])
            assert relerr(gd, pb.a["g_dead"][: gd.shape[0]]) <= VTOL
        return out
    finally:
        if tr.built:
            tr.destroy()


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 33, 100, 1000, 5000])
def test_small_clouds(ctx, port, n):
    run_pair(ctx, port, cases.cloud(n, "gauss", "same", seed=n))


def test_empty(ctx, port):
    from vvflow_b200 import vvhd
    S = vvhd.Space(ctx=ctx)
    S.VortexList = np.zeros((0, 3))
    tr = vvhd.TSortedTree(S, 8, 0.0)
    tr.build()
    assert ctx.tree_counts()[:2] == (1, 1)
    vvhd.MEpsilonFast(S, tr).CalcEpsilonFast(True)
    vvhd.MConvectiveFast(S, tr).process_all_lists()
    vvhd.MDiffusiveFast(S, tr).process_vort_list()
    tr.destroy()
    assert vvhd.MFlowmove(S).move_and_clean(True) == 0
    assert S.VortexList.shape == (0, 6)


def test_gauss_same_sign_100k(ctx, port):
    out = run_pair(ctx, port, cases.cloud(100_000, "gauss", "same", seed=7), re=1000.0, dt=0.005)
    assert out["merged"] == 0


def test_uniform_same_sign_50k(ctx, port):
    run_pair(ctx, port, cases.cloud(50_000, "uniform", "same", seed=8), re=1000.0, dt=0.005)


def test_duplicates_and_lines(ctx, port):
    """coincident particles, particles on a line (degenerate boxes), zero-circulation particles"""
    rng = np.random.default_rng(5)
    a = cases.cloud(3000, "gauss", "same", seed=5)
    a[:500, :2] = a[500:1000, :2]           # exact duplicates
    a[1000:1500, 1] = 0.25                  # a horizontal line
    a[1500:1600, 2] = 0.0                   # g == 0: skipped as source and as target
    a[1600:1700, :2] = 1.5                  # 100 coincident
    run_pair(ctx, port, a)
    b = np.zeros((40, 3)); b[:, 0] = rng.uniform(0, 1, 40); b[:, 2] = 1.0  # all on y = 0
    run_pair(ctx, port, b)


@pytest.mark.parametrize("sign", ["same", "mixed"])
def test_big_leaves(ctx, port, sign):
    """minNodeSize stops the subdivision early (TSortedTree.cpp:45-47): leaves of ~60 particles, so that the
    near-field kernels run several 15-target passes per leaf and expand source leaves in pieces"""
    xyg = cases.cloud(20000, "uniform", sign, seed=77)
    out = run_pair(ctx, port, xyg, tree=(0.06, 1e300), merge=(sign == "mixed"))
    assert out["leaves"] < 20000 // 30


@pytest.mark.parametrize("n,seed", [(200, 1), (2000, 2), (20000, 3)])
def test_mixed_sign_merge(ctx, port, n, seed):
    """BASELINE config 5b: order-dependent merging, ~23 % of the particles merge"""
    out = run_pair(ctx, port, cases.cloud(n, "gauss", "mixed", seed=seed))
    assert out["merged"] > 0


def test_merge_disabled(ctx, port):
    out = run_pair(ctx, port, cases.cloud(3000, "gauss", "mixed", seed=9), merge=False, re=float("inf"))
    assert out["merged"] == 0


@pytest.mark.parametrize("n,sign", [(300, "same"), (8000, "mixed"), (30000, "mixed")])
def test_cylinder(ctx, port, n, sign):
    """BASELINE config 1 geometry: cylinder R=0.5 with 350 segments, particles in a shell around it
    (wall epsilon restriction, wall merge criterion, segment diffusion + fric, in-body removal)."""
    body = cases.cylinder(0.5, 350)
    out = run_pair(ctx, port, cases.around_cylinder(n, sign=sign, seed=n), bodies=[body])
    assert out["merged"] >= 0


def test_two_cylinders_moving(ctx, port):
    """BASELINE config 4 geometry: two tandem cylinders, one with speed_slae != 0 and slip segments
    (body_list_influence incl. the linear-source term)"""
    b1 = cases.cylinder(0.5, 200, 0.0, 0.0)
    b2 = cases.cylinder(0.5, 200, 2.0, 0.0)
    b2.speed_slae = np.array([0.1, -0.05, 0.2])
    b2.slip[:20] = 1
    b2.g[:20] = 0.01
    b2.axis = np.array([2.0, 0.0])
    rng = np.random.default_rng(11)
    n = 6000
    xyg = np.zeros((n, 3))
    xyg[:, 0] = rng.uniform(-1, 3.5, n); xyg[:, 1] = rng.uniform(-1.2, 1.2, n)
    xyg[:, 2] = rng.uniform(-1, 1, n) / n
    # a few particles hugging a segment of the moving body (exercise the near branch)
    xyg[:50, 0] = 2.0 + 0.5005 * np.cos(np.linspace(0, 1, 50)); xyg[:50, 1] = 0.5005 * np.sin(np.linspace(0, 1, 50))
    run_pair(ctx, port, xyg, bodies=[b1, b2])


def test_sinks_process_all_lists(ctx, port):
    """sink_list_influence inside process_all_lists (MConvectiveFast.cpp:82,153-170), body-free and
    with bodies; strengths below and above 1 (the reference truncates the strength to an integer, :165)"""
    sinks = np.array([[0.3, 0.2, 0.2], [-0.5, 0.1, -1.5], [0.0, -0.4, 2.7], [1.1, 0.9, -0.99]])
    run_pair(ctx, port, cases.cloud(20000, "gauss", "mixed", seed=44), sinks=sinks, inf=(1.0, 0.25))
    run_pair(ctx, port, cases.around_cylinder(8000, sign="mixed", seed=45), bodies=[cases.cylinder(0.5, 350)],
             sinks=sinks + np.array([1.5, 0.0, 0.0]))


def test_sinks_golden(ctx):
    """the same path against the compiled reference's own values (tests/golden/sinks_1500.npz)"""
    from vvflow_b200 import vvhd
    d = cases.golden("sinks_1500")
    re, dt, ivx, ivy = d["params"]
    S = vvhd.Space(ctx=ctx)
    S.VortexList = d["xyg"]
    S.SourceList = d["sinks"]
    S.re, S.dt, S.inf_vx, S.inf_vy = re, dt, ivx, ivy
    tr = vvhd.TSortedTree(S, 8, 0.0)
    try:
        tr.build()
        vvhd.MEpsilonFast(S, tr).CalcEpsilonFast(False)
        conv = vvhd.MConvectiveFast(S, tr)
        check_close(conv.velocity(d["pts"]), d["vel_at_pts"], VTOL, "velocity(p) with sinks")
        conv.process_all_lists()
        a = S.VortexList
        assert same(a[:, [0, 1, 2, 5]], d["after_conv"][:, [0, 1, 2, 5]])
        check_close(a[:, 3:5], d["after_conv"][:, 3:5], VTOL, "process_all_lists with sinks")
    finally:
        if tr.built:
            tr.destroy()


def _points_for(xyg, rng, n=2000):
    lo, hi = xyg[:, :2].min(0), xyg[:, :2].max(0)
    inside = rng.uniform(lo, hi, (n, 2))
    on = xyg[rng.integers(0, xyg.shape[0], 50), :2]          # dr = 0 with a source (eps keeps it finite)
    far = rng.uniform(lo - 3 * (hi - lo), hi + 3 * (hi - lo), (200, 2))
    return np.concatenate([inside, on, far])


@pytest.mark.parametrize("case", ["cloud", "cylinder", "moving_bodies_sinks"])
def test_velocity_at_points(ctx, port, case):
    """MConvectiveFast::velocity(p) (MConvectiveFast.cpp:20-34) against the oracle, which is
    pinned bit-exact to the compiled reference for this function (tests/test_oracle_port.py). 1e-10 norm-wise."""
    from vvflow_b200 import vvhd
    rng = np.random.default_rng(31)
    bodies, sinks = [], None
    if case == "cloud":
        xyg = cases.cloud(30000, "gauss", "mixed", seed=12)
    elif case == "cylinder":
        xyg = cases.around_cylinder(8000, sign="mixed", seed=13)
        bodies = [cases.cylinder(0.5, 350)]
    else:
        b1 = cases.cylinder(0.5, 200, 0.0, 0.0)
        b2 = cases.cylinder(0.5, 200, 2.0, 0.0)
        b2.speed_slae = np.array([0.1, -0.05, 0.2]); b2.slip[:20] = 1; b2.g[:20] = 0.01; b2.axis = np.array([2.0, 0.0])
        bodies = [b1, b2]
        xyg = np.zeros((6000, 3))
        xyg[:, 0] = rng.uniform(-1, 3.5, 6000); xyg[:, 1] = rng.uniform(-1.2, 1.2, 6000)
        xyg[:, 2] = rng.uniform(-1, 1, 6000) / 6000
        sinks = np.array([[0.3, 1.5, 0.2], [-2.0, 0.1, -0.1]])
    mn, mx = cases.tree_params(bodies)
    pb = cases.port_bodies(port, bodies)
    P = port.Port(xyg=xyg, bodies=pb)
    S = vvhd.Space(ctx=ctx)
    S.VortexList = xyg
    S.BodyList = bodies
    S.re, S.dt, S.inf_vx, S.inf_vy = 600.0, 0.05, 1.0, 0.25
    if sinks is not None:
        S.SourceList = sinks
    tr = vvhd.TSortedTree(S, 8, mn, mx)
    eps, conv = vvhd.MEpsilonFast(S, tr), vvhd.MConvectiveFast(S, tr)
    with pytest.raises(RuntimeError):
        conv.velocity((0.0, 0.0))                       # findNode on an unbuilt tree throws (TSortedTree.cpp:286-288)
    try:
        P.tree_build(8, mn, mx); tr.build()
        eps.CalcEpsilonFast(True)
        assert P.epsilon(True) == eps.Merged()
        pts = _points_for(xyg, rng)
        want = P.velocity_at(pts, 1.0, 0.25, 0.05, sinks)
        got = conv.velocity(pts)
        ok = np.isfinite(want).all(axis=1)
        assert ok.sum() >= pts.shape[0] - 2
        check_close(got[ok], want[ok], VTOL, "velocity(p)")
        one = conv.velocity(pts[7])
        assert one.shape == (2,) and np.array_equal(one, got[7])
        assert conv.velocity(np.zeros((0, 2))).shape == (0, 2)
    finally:
        if tr.built:
            tr.destroy()
        P.tree_destroy()


@pytest.mark.parametrize("with_body", [False, True])
def test_eps2h_h2_at_points(ctx, port, with_body):
    """static MEpsilonFast::eps2h / h2 (MEpsilonFast.cpp:66-107) of findNode(p): order-free, so bit-exact"""
    from vvflow_b200 import vvhd
    rng = np.random.default_rng(41)
    bodies = [cases.cylinder(0.5, 350)] if with_body else []
    xyg = cases.around_cylinder(8000, sign="mixed", seed=23) if with_body else cases.cloud(30000, "gauss", "mixed", seed=22)
    xyg[:40, :2] = xyg[40:80, :2]            # duplicates: zero distances are skipped, ties count twice
    mn, mx = cases.tree_params(bodies)
    P = port.Port(xyg=xyg, bodies=cases.port_bodies(port, bodies))
    S = vvhd.Space(ctx=ctx)
    S.VortexList = xyg
    S.BodyList = bodies
    tr = vvhd.TSortedTree(S, 8, mn, mx)
    eps = vvhd.MEpsilonFast(S, tr)
    try:
        P.tree_build(8, mn, mx); tr.build()
        pts = _points_for(xyg, rng)
        want = P.eps2h_h2_at(pts)
        assert same(eps.eps2h(pts), want[:, 0]), "eps2h differs"
        assert same(eps.h2(pts), want[:, 1]), "h2 differs"
        assert with_body == bool(np.isfinite(want[:, 1]).any())
        assert eps.eps2h(pts[3]) == want[3, 0]
    finally:
        if tr.built:
            tr.destroy()
        P.tree_destroy()


@pytest.mark.parametrize("spread", [0.05, 0.3])
def test_node_influence(ctx, port, spread):
    """MConvectiveFast::NodeInfluence (MConvectiveFast.cpp:398-418) for every segment, the vortex
    term of the SLAE right-hand side. Oracle pinned bit-exact to the compiled reference; 1e-10 norm-wise here
    (log / sqrt of the device library, order of summation)."""
    from vvflow_b200 import vvhd
    bodies = [cases.cylinder(0.5, 350), cases.cylinder(0.3, 120, 1.6, 0.2)]
    xyg = cases.around_cylinder(9000, sign="mixed", seed=37, spread=spread)
    mn, mx = cases.tree_params(bodies)
    P = port.Port(xyg=xyg, bodies=cases.port_bodies(port, bodies))
    S = vvhd.Space(ctx=ctx)
    S.VortexList = xyg
    S.BodyList = bodies
    tr = vvhd.TSortedTree(S, 8, mn, mx)
    eps, conv = vvhd.MEpsilonFast(S, tr), vvhd.MConvectiveFast(S, tr)
    with pytest.raises(RuntimeError):
        conv.NodeInfluence()
    try:
        P.tree_build(8, mn, mx); tr.build()
        eps.CalcEpsilonFast(False); P.epsilon(False)          # core radii
        want = P.node_influence()
        got = conv.NodeInfluence()
        assert got.shape == want.shape == (470,)
        assert np.isfinite(want).all() and np.abs(want).max() > 0
        assert relerr(got, want) <= VTOL, relerr(got, want)
    finally:
        if tr.built:
            tr.destroy()
        P.tree_destroy()


def _gpu_step(ctx, dt=0.05, re=600.0):
    ctx.tree_build(8, 0.0)
    ctx.epsilon(True)
    ctx.convective(1.0, 0.0, dt)
    ctx.diffusive(re, want_fric=False)
    ctx.tree_destroy()
    ctx.move_and_clean(dt)


def test_append_particles_resident_loop(ctx):
    from vvflow_b200 import capi
    """SURVEY 8(f) row 2: the list stays on the device between steps and newly shed vortices are appended
    (vvgpu_append_particles) instead of re-uploading everything. The step on (resident survivors + appended) must be
    bit-identical to the step on the same list uploaded afresh, and the caller-order indices must continue."""
    a = np.zeros((20000, 6)); a[:, :3] = cases.cloud(20000, "gauss", "mixed", seed=51)
    b = np.zeros((3000, 6)); b[:, :3] = cases.cloud(3000, "uniform", "mixed", seed=52); b[:, 5] = 7.5   # carries an _1_eps
    ctx.set_particles(a)
    _gpu_step(ctx)
    surv = ctx.get_particles()
    ctx.append_particles(b)
    both = ctx.get_particles()
    assert both.shape[0] == surv.shape[0] + 3000
    assert same(both[: surv.shape[0]], surv) and same(both[surv.shape[0]:], b)
    perm = ctx.get_permutation()
    assert np.array_equal(perm[surv.shape[0]:], 20000 + np.arange(3000))
    ctx.append_particles(np.zeros((0, 6)))              # appending nothing is a no-op
    assert ctx.n == both.shape[0]
    _gpu_step(ctx)
    resident = ctx.get_particles()
    ctx.set_particles(both)
    _gpu_step(ctx)
    fresh = ctx.get_particles()
    assert resident.shape == fresh.shape and same(resident, fresh)
    ctx.tree_build(8, 0.0)
    with pytest.raises(capi.VVGpuError):
        ctx.append_particles(b)                         # the tree holds positions into the list
    ctx.tree_destroy()


@pytest.mark.parametrize("with_body", [False, True])
def test_vorticity_raster(ctx, port, with_body):
    """XVorticity::evaluate (XVorticity.cpp:26-97). The oracle's raster is pinned to the compiled
    reference (bit-exact after the float rounding of XField::map); here 1e-10 norm-wise in double, and the resident
    list must come back untouched (the reference works on a copy of the Space)."""
    from vvflow_b200 import vvhd
    if with_body:
        bodies = [cases.cylinder(0.5, 350)]
        bodies[0].g[:] = np.random.default_rng(81).uniform(-1, 1, 350) * 1e-3     # attached vortices to shed
        xyg = cases.around_cylinder(8000, sign="mixed", seed=82)
        grid = (-1.0, -1.0, 0.02, 100, 100, 1.5)
    else:
        bodies = []
        xyg = cases.cloud(20000, "gauss", "mixed", seed=83)
        grid = (-2.5, -2.5, 0.05, 100, 100, 2.0)
    S = vvhd.Space(ctx=ctx)
    S.VortexList = xyg
    S.BodyList = bodies
    f = vvhd.XVorticity(S, *grid[:5])
    with pytest.raises(ValueError):
        f.evaluate()                                   # eps_mult must be positive
    f.eps_mult = grid[5]
    f.evaluate()
    assert same(ctx.get_particles()[:, :3], xyg)       # the Space's list is back on the device, in its own order
    # oracle on the same post-shed list
    shed = []
    for b in bodies:
        keep = (np.abs(b.g) >= 1e-10) & (b.slip == 0)
        shed.append(np.stack([b.corner[keep, 0] - b.dl[keep, 1] * 1e-4, b.corner[keep, 1] + b.dl[keep, 0] * 1e-4, b.g[keep]], 1))
    P = port.Port(xyg=np.concatenate([xyg] + shed), bodies=cases.port_bodies(port, bodies))
    want = P.vorticity_raster(*[np.float32(v) for v in grid[:3]], grid[3], grid[4], grid[5], S.average_segment_length())
    got = ctx_raster = f.map
    assert got.shape == want.shape == (100, 100)
    assert np.abs(want).max() > 0 and (not with_body or (want == 0).sum() > 100)
    assert np.array_equal(got == 0, want.astype(np.float32) == 0)                 # the same points are inside the body
    assert relerr(got.astype(np.float64), want) <= 2e-7                            # float32 map
    # the double values behind the map
    S.ctx.set_particles(np.concatenate([np.concatenate([xyg, np.zeros((xyg.shape[0], 3))], 1)] +
                                       [np.concatenate([s_, np.zeros((s_.shape[0], 3))], 1) for s_ in shed]))
    dbl = ctx.vorticity_raster(float(np.float32(grid[0])), float(np.float32(grid[1])), float(np.float32(grid[2])), grid[3], grid[4],
                               grid[5], S.average_segment_length())
    assert relerr(dbl, want) <= VTOL, relerr(dbl, want)


@pytest.mark.parametrize("case", ["cloud", "moving_cylinder_b", "moving_cylinder_s"])
def test_pressure_raster(ctx, port, case):
    """XPressure::evaluate (XPressure.cpp:32-146). The oracle's raster is pinned to the compiled
    reference (tests/test_oracle_port.py); here 1e-10 norm-wise on the double values (the direct sum over all vortices
    runs in another order) and the same zeros inside the body; the resident list comes back untouched."""
    from vvflow_b200 import vvhd
    if case == "cloud":
        bodies, xyg, grid, frame = [], cases.cloud(20000, "uniform", "mixed", seed=91), (-0.5, -0.5, 0.02, 100, 100), "o"
    else:
        bodies = [cases.cylinder(0.5, 350)]
        rng = np.random.default_rng(92)
        bodies[0].g[:] = rng.uniform(-1, 1, 350) * 1e-3          # attached vortices to shed
        bodies[0].gsum[:] = rng.uniform(-1, 1, 350) * 1e-3
        bodies[0].speed_slae = np.array([0.1, -0.05, 0.2])       # the first addend of :115-121 is live
        xyg = cases.around_cylinder(12000, sign="mixed", seed=93)
        grid, frame = (-1.0, -1.0, 0.02, 100, 100), case[-1]
    S = vvhd.Space(ctx=ctx)
    S.VortexList = xyg
    S.BodyList = bodies
    S.re, S.dt, S.inf_vx, S.inf_vy = 600.0, 0.05, 1.0, 0.25
    f = vvhd.XPressure(S, *grid)
    with pytest.raises(ValueError):
        f.evaluate()                                   # eps_mult must be positive
    f.eps_mult = 1.0
    f.ref_frame = "x"
    with pytest.raises(ValueError):
        f.evaluate()                                   # bad ref_frame
    f.ref_frame = frame
    f.evaluate()
    assert same(ctx.get_particles()[:, :3], xyg)       # the Space's list is back on the device, in its own order
    shed = []
    for b in bodies:
        keep = (np.abs(b.g) >= 1e-10) & (b.slip == 0)
        shed.append(np.stack([b.corner[keep, 0] - b.dl[keep, 1] * 1e-4, b.corner[keep, 1] + b.dl[keep, 0] * 1e-4, b.g[keep]], 1))
    pb = cases.port_bodies(port, bodies)
    if pb is not None:
        pb.a["gsum"][:] = np.concatenate([b.gsum + b.g for b in bodies])
    P = port.Port(xyg=np.concatenate([xyg] + shed), bodies=pb)
    ref = {"s": None, "o": (0.0, 0.0), "b": (0.1, -0.05)}[frame]
    want = P.pressure_raster(*[np.float32(v) for v in grid[:3]], grid[3], grid[4], S.average_segment_length(), 600.0, 0.05, 1.0, 0.25,
                             ref_speed=ref)
    got = f.map
    assert got.shape == want.shape == (100, 100) and np.isfinite(want).all() and np.abs(want).max() > 0
    assert np.array_equal(got == 0, want.astype(np.float32) == 0) and (not bodies or (want == 0).sum() > 100)
    assert relerr(got.astype(np.float64), want) <= 2e-7                            # float32 map
    # the double values behind the map
    rec = np.concatenate([xyg] + shed)
    S.ctx.set_particles(np.concatenate([rec, np.zeros((rec.shape[0], 3))], 1))
    dbl = ctx.pressure_raster(float(np.float32(grid[0])), float(np.float32(grid[1])), float(np.float32(grid[2])), grid[3], grid[4],
                              S.average_segment_length(), 600.0, 0.05, 1.0, 0.25,
                              np.concatenate([b.gsum + b.g for b in bodies]) if bodies else None, None, ref)
    assert relerr(dbl, want) <= VTOL, relerr(dbl, want)


def test_awkward_small_inputs(ctx, port):
    """the sweep of tests/test_oracle_port.py::test_port_matches_reference_on_random_small_inputs on the CUDA path:
    one or two particles, exact duplicates, points on a line, a lattice (ties in every comparison), zero
    circulations, particles inside a body"""
    rng = np.random.default_rng(2024)
    for trial in range(30):
        n = int(rng.choice([1, 2, 3, 15, 16, 17, 31, 64, 200, 400]))
        kind = trial % 5
        if kind == 0:
            xy = rng.standard_normal((n, 2))
        elif kind == 1:
            xy = rng.uniform(-1, 1, (n, 2)); xy[n // 2:] = xy[: n - n // 2]
        elif kind == 2:
            xy = np.stack([np.linspace(-1, 1, n), np.zeros(n)], 1)
        elif kind == 3:
            k = int(np.ceil(np.sqrt(n)))
            xy = np.stack(np.meshgrid(np.arange(k), np.arange(k)), -1).reshape(-1, 2)[:n] * 0.125 - 0.3
        else:
            rad = 0.5 + np.abs(rng.standard_normal(n)) * 0.2; th = rng.uniform(0, 2 * np.pi, n)
            xy = np.stack([rad * np.cos(th), rad * np.sin(th)], 1); xy[: n // 8] *= 0.3
        g = rng.uniform(-1, 1, n) / n
        g[rng.uniform(size=n) < 0.1] = 0.0
        xyg = np.concatenate([xy.astype(np.float64), g[:, None]], 1)
        bodies = [cases.cylinder(0.5, 60)] if (kind == 4 or trial % 7 == 0) else []
        run_pair(ctx, port, xyg, bodies=bodies, merge=(trial % 3 != 0))


def test_against_reference_build(ctx, ref):
    """same comparison directly against the reference's own compiled code, where it travelled"""
    xyg = cases.cloud(20000, "gauss", "mixed", seed=21)
    r = ref.Ref(re=600, dt=0.05, inf_vx=1.0)
    r.set_list(xyg)
    mn, mx = r.tree_params(8)
    r.tree_build()
    from vvflow_b200 import vvhd
    S = vvhd.Space(ctx=ctx)
    S.VortexList = xyg
    S.re, S.dt, S.inf_vx = 600.0, 0.05, 1.0
    tr = vvhd.TSortedTree(S, 8, mn, mx)
    tr.build()
    assert same(S.VortexList[:, :3], r.get_list48()[:, :3])
    m = r.epsilon(True)
    e = vvhd.MEpsilonFast(S, tr); e.CalcEpsilonFast(True)
    assert e.Merged() == m
    assert same(S.VortexList[:, [0, 1, 2, 5]], r.get_list48()[:, [0, 1, 2, 5]])
    r.convective(); vvhd.MConvectiveFast(S, tr).process_all_lists()
    r.diffusive(); vvhd.MDiffusiveFast(S, tr).process_vort_list()
    check_close(S.VortexList[:, 3:5], r.get_list48()[:, 3:5], VTOL, "velocity vs compiled reference")
    r.tree_destroy(); tr.destroy()
    r.move_and_clean(True); vvhd.MFlowmove(S).move_and_clean(True)
    assert S.VortexList.shape == r.get_list48().shape
    check_close(S.VortexList[:, :2], r.get_list48()[:, :2], VTOL, "positions vs compiled reference")


def test_full_step_1M_against_oracle(ctx, port):
    """the benchmarked input itself (bench.py, BASELINE configs[1], N = 1M) through one full step against the oracle:
    tree (depth 21, 96k leaves, fringe and multi-unit groups that only occur at scale), permutation and epsilon
    BIT-EXACT, velocities and advected positions 1e-10 norm-wise and element-wise. ~1 min of CPU for the oracle."""
    import bench
    w = bench.make_workload("lamb", 1_000_000)
    out = run_pair(ctx, port, w["rec"][:, :3], re=w["re"], dt=w["dt"], inf=w["inf"], lists=False)
    assert out["merged"] == 0 and out["leaves"] > 90_000


def test_cylinder_200k_with_merges(ctx, port):
    """a body case at scale: 200k mixed-sign particles around the 350-segment cylinder (merge fixed point over several
    rounds, wall passes, segment diffusion + fric, in-body removal) against the oracle"""
    import bench
    w = bench.make_workload("cyl", 200_000)
    from vvflow_b200 import vvhd
    out = run_pair(ctx, port, w["rec"][:, :3], bodies=[vvhd.TBody(b) for b in w["bodies"]], re=w["re"], dt=w["dt"],
                   inf=w["inf"], lists=False)
    assert out["merged"] > 1000


def test_conv_tma_variant():
    """the build of K4 that stages the source leaves with TMA bulk copies (lib/libvvgpu_tma.so, -DVV_CV_TMA=1; measured
    slower than the default, kept as the evidence behind that choice): same convective velocities, 1e-10, on a cloud and
    a cylinder case, in a fresh process so that the other library is the one loaded"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "vvflow_b200", "lib", "libvvgpu_tma.so")
    if not os.path.exists(lib):
        pytest.skip("libvvgpu_tma.so not built (vvflow_b200.build.build(variants=True))")
    code = """
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import cases
from oracle import pyport
from vvflow_b200 import capi, vvhd
ctx = capi.Context(0)
for xyg, bodies in ((cases.cloud(30000, "gauss", "mixed", seed=5), []),
                    (cases.around_cylinder(20000, sign="mixed", seed=6), [cases.cylinder(0.5, 350)])):
    mn, mx = cases.tree_params(bodies)
    P = pyport.Port(xyg=xyg, bodies=cases.port_bodies(pyport, bodies))
    S = vvhd.Space(ctx=ctx); S.VortexList = xyg; S.BodyList = bodies; S.re, S.dt, S.inf_vx = 600., 0.05, 1.
    tr = vvhd.TSortedTree(S, 8, mn, mx)
    P.tree_build(8, mn, mx); tr.build()
    e = vvhd.MEpsilonFast(S, tr); e.CalcEpsilonFast(True)
    assert P.epsilon(True) == e.Merged()
    P.convective(1.0, 0.0, 0.05); vvhd.MConvectiveFast(S, tr).process_all_lists()
    a, b = S.VortexList, P.rec48()
    assert cases.same(a[:, [0, 1, 2, 5]], b[:, [0, 1, 2, 5]])
    cases.check_close(a[:, 3:5], b[:, 3:5], 1e-10, "convective velocity, TMA variant")
    tr.destroy(); P.tree_destroy()
print("TMA-OK")
""" % (root, os.path.join(root, "tests"))
    env = dict(os.environ, VVGPU_LIB=lib)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "TMA-OK" in p.stdout, p.stdout[-1500:] + p.stderr[-1500:]


def test_convective_without_epsilon(ctx, port):
    """_1_eps == 0 on every particle (a list that never went through CalcEpsilonFast): the reference's near field is
    g / (dr^2 + inf) = 0 per pair and only the far field is left; no NaN from the fast reciprocal"""
    from vvflow_b200 import vvhd
    xyg = cases.cloud(3000, "gauss", "mixed", seed=77)
    P = port.Port(xyg=xyg, bodies=cases.port_bodies(port, []))
    S = vvhd.Space(ctx=ctx)
    S.VortexList = xyg
    S.dt, S.inf_vx, S.inf_vy = 0.05, 1.0, 0.0
    tr = vvhd.TSortedTree(S, 8, 0.0)
    conv = vvhd.MConvectiveFast(S, tr)
    P.tree_build(8, 0.0, float(np.finfo(np.float64).max))
    tr.build()
    try:
        P.convective(1.0, 0.0, 0.05)
        conv.process_all_lists()
        a, b = S.VortexList, P.rec48()
        assert np.isfinite(a[:, 3:5]).all()
        check_close(a[:, 3:5], b[:, 3:5], VTOL, "convective velocity with eps = inf")
    finally:
        tr.destroy()


def test_error_behaviour(ctx):
    """call-order errors mirror the reference (TSortedTree.cpp:234,286-288; MFlowmove.cpp:20-22)"""
    from vvflow_b200 import capi, vvhd
    S = vvhd.Space(ctx=ctx)
    S.VortexList = cases.cloud(100)
    tr = vvhd.TSortedTree(S, 8, 0.0)
    with pytest.raises(capi.VVGpuError):
        ctx.epsilon(True)           # tree is not built
    with pytest.raises(ValueError):
        tr.findNode((0.0, 0.0))
    tr.build()
    with pytest.raises(capi.VVGpuError):
        ctx.tree_build()            # already built
    with pytest.raises(capi.VVGpuError):
        ctx.set_particles(np.zeros((3, 6)))
    assert tr.findNode((0.0, 0.0)) >= 0
    tr.destroy()
    with pytest.raises(ValueError):
        vvhd.MFlowmove(S).move_and_clean(True, collision=None)


def test_properties_at_scale(ctx):
    """BASELINE config 2 size (N = 1M): size-independent properties instead of a CPU oracle run.
    permutation is a bijection; leaves tile the array; every particle lies inside its leaf box;
    sum of g is conserved; velocity of a same-sign blob has the right circulation sense."""
    from vvflow_b200 import vvhd
    n = 1_000_000
    xyg = cases.cloud(n, "gauss", "equal", seed=12345)
    S = vvhd.Space(ctx=ctx)
    S.VortexList = xyg
    S.re, S.dt, S.inf_vx = 1000.0, 0.005, 0.0
    tr = vvhd.TSortedTree(S, 8, 0.0)
    tr.build()
    perm = ctx.get_permutation()
    assert np.array_equal(np.sort(perm), np.arange(n))
    v = S.VortexList
    assert np.array_equal(v[:, :3], xyg[perm])
    leaves = tr.getBottomNodes()
    first, last = leaves[:, 4].astype(np.int64), leaves[:, 5].astype(np.int64)
    assert first[0] == 0 and last[-1] == n and np.array_equal(first[1:], last[:-1])
    assert np.max(last - first) < 16
    lid = np.repeat(np.arange(leaves.shape[0]), last - first)
    # inside the (tight) leaf box, up to the rounding of the box centre (bl+tr)*0.5
    assert np.all(np.abs(v[:, 0] - leaves[lid, 0]) <= 0.5 * leaves[lid, 3] + 4e-16 * (1 + np.abs(leaves[lid, 0])))
    assert np.all(np.abs(v[:, 1] - leaves[lid, 1]) <= 0.5 * leaves[lid, 2] + 4e-16 * (1 + np.abs(leaves[lid, 1])))
    e = vvhd.MEpsilonFast(S, tr); e.CalcEpsilonFast(True)
    assert e.Merged() == 0
    vvhd.MConvectiveFast(S, tr).process_all_lists()
    v = S.VortexList
    assert np.all(np.isfinite(v)) and np.all(v[:, 5] > 0)
    # Lamb-Oseen: azimuthal velocity u_theta(r) = G/(2 pi r) * (1 - exp(-r^2/2)), total G = 1
    r = np.hypot(v[:, 0], v[:, 1])
    ut = (-v[:, 1] * v[:, 3] + v[:, 0] * v[:, 4]) / r
    sel = (r > 0.5) & (r < 2.0)
    exact = (1 - np.exp(-r[sel] ** 2 / 2)) / (2 * np.pi * r[sel])
    assert abs(np.mean(ut[sel] / exact) - 1) < 0.02
    tr.destroy()
    vvhd.MFlowmove(S).move_and_clean(True)
    assert S.VortexList.shape[0] == n and abs(np.sum(S.VortexList[:, 2]) - 1.0) < 1e-9


@pytest.mark.parametrize("name", cases.GOLDEN)
def test_golden_fixtures(ctx, name):
    """the CUDA path against the vectors generated from the reference's own compiled code"""
    from vvflow_b200 import vvhd
    d = cases.golden(name)
    bodies = cases.golden_bodies(d)
    far, mn, mx = d["tree_params"]
    re, dt, ivx, ivy = d["params"]
    S = vvhd.Space(ctx=ctx)
    S.VortexList = d["in48"]
    S.BodyList = bodies
    S.re, S.dt, S.inf_vx, S.inf_vy = re, dt, ivx, ivy
    tr = vvhd.TSortedTree(S, int(far), mn, mx)
    tr.build()
    try:
        dbl, idx, nl = ctx.tree_export()
        assert nl == d["n_leaves"][0]
        assert same(dbl, d["tree_dbl"])
        assert same(idx[:, [0, 1, 2, 3, 4, 5]], d["tree_idx"][:, [0, 1, 6, 7, 8, 9]])
        for a, k in zip(ctx.tree_lists(), ("near_ptr", "near_idx", "far_ptr", "far_idx")):
            assert same(a, d[k]), k
        assert same(S.VortexList, d["after_build"])
        assert ctx.count_interactions() == (d["interactions"][0], d["interactions"][1])
        e = vvhd.MEpsilonFast(S, tr); e.CalcEpsilonFast(True)
        assert e.Merged() == d["merged"][0]
        assert same(S.VortexList, d["after_eps"])
        # SURVEY 8(f) rows 1, 4 against the reference's own values
        conv = vvhd.MConvectiveFast(S, tr)
        ok = np.isfinite(d["vel_at_pts"]).all(axis=1)
        check_close(conv.velocity(d["pts"])[ok], d["vel_at_pts"][ok], VTOL, "velocity(p) vs golden")
        assert same(e.eps2h(d["pts"]), d["eps2h_h2_at_pts"][:, 0]) and same(e.h2(d["pts"]), d["eps2h_h2_at_pts"][:, 1])
        if bodies:
            assert relerr(conv.NodeInfluence(), d["node_influence"]) <= VTOL
        vvhd.MConvectiveFast(S, tr).process_all_lists()
        check_close(S.VortexList[:, 3:5], d["after_conv"][:, 3:5], VTOL, "convective vs golden")
        vvhd.MDiffusiveFast(S, tr).process_vort_list()
        check_close(S.VortexList[:, 3:5], d["after_diff"][:, 3:5], VTOL, "diffusive vs golden")
        if bodies:
            fr = np.concatenate([b.fric for b in bodies])
            assert relerr(fr, d["seg_after_diff"][:, 8] - d["seg_in"][:, 8]) <= 1e-9
    finally:
        tr.destroy()
    cleaned = vvhd.MFlowmove(S).move_and_clean(True)
    assert cleaned == d["cleaned"][0]
    a = S.VortexList
    assert a.shape == d["after_move"].shape
    assert same(a[:, 2], d["after_move"][:, 2])
    check_close(a[:, :2], d["after_move"][:, :2], VTOL, "positions vs golden")
    if bodies:
        gs = np.concatenate([b.gsum for b in bodies])
        assert relerr(gs, d["seg_after_move"][:, 7] - d["seg_in"][:, 7]) <= 1e-9
        assert relerr(bodies[0].fdt_dead, d["body_after_move"][0, 13:16] - d["body_in"][0, 13:16]) <= 1e-9
