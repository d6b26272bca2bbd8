"""Scratch: hot-path passes of the cylinder workload (for ncu captures). usage: prof_cyl.py N reps"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from vvflow_b200 import capi, vvhd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = bench.make_workload("cyl", n)
ctx = capi.Context(0)
S = vvhd.Space(ctx=ctx)
S.BodyList = [vvhd.TBody(b) for b in w["bodies"]]
ctx.set_bodies(*S._pack_bodies())
dl = S.average_segment_length()
for _ in range(reps):
    ctx.set_particles(w["rec"])
    ctx.tree_build(8, dl * 5, dl * 100)
    m = ctx.epsilon(True)
    ctx.convective(1.0, 0.0, w["dt"])
    ctx.diffusive(w["re"], want_fric=True)
    ctx.tree_destroy()
    ctx.move_and_clean(w["dt"])
ctx.synchronize()
print(ctx.phase_times(), "merged", m, "rounds", ctx.merge_rounds())
