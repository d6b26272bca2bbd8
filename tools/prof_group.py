"""Scratch: P ranks of an in-process group on ONE device, for an ncu launch list of the sharded step (ncu serialises
the kernels, so each rank's kernel times are those it would see on its own GPU; the copies between ranks are not).
usage: prof_group.py N P [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from vvflow_b200 import capi, multigpu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
xyg = cases.cloud(n, "gauss", "equal", seed=12345)
ctxs = capi.group_create([0] * P)


def work(r, ctx):
    for _ in range(reps):
        ctx.set_particles_xyg(xyg)
        ctx.tree_build(8, 0.0)
        ctx.epsilon(True)
        ctx.convective(1.0, 0.0, 0.005)
        ctx.diffusive(1000.0, want_fric=False)
        ctx.tree_destroy()
        ctx.move_and_clean(0.005)
    ctx.synchronize()
    return ctx.phase_times()


for r, t in enumerate(multigpu.run_group(ctxs, work)):
    print(r, {k: round(v, 3) for k, v in t[0].items()})
