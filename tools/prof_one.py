"""Scratch: one hot-path pass at a given N (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from vvflow_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = capi.Context(0)
xyg = cases.cloud(n, "gauss", "equal", seed=12345)
for _ in range(reps):
    ctx.set_particles_xyg(xyg)
    ctx.tree_build(8, 0.0)
    ctx.epsilon(True)
    ctx.convective(1.0, 0.0, 0.005)
    ctx.diffusive(1000.0, want_fric=False)
    ctx.tree_destroy()
    ctx.move_and_clean(0.005)
ctx.synchronize()
print(ctx.phase_times())
