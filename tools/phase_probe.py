"""Scratch: phase times (CUDA events inside the library) of a few steps at N, best of reps. VVGPU_LIB picks the build."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from vvflow_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = capi.Context(0)
xyg = cases.cloud(n, "gauss", "equal", seed=12345)
best = None
for _ in range(reps):
    ctx.set_particles_xyg(xyg)
    ctx.tree_build(8, 0.0)
    ctx.epsilon(True)
    ctx.convective(1.0, 0.0, 0.005)
    ctx.diffusive(1000.0, want_fric=False)
    ctx.tree_destroy()
    ctx.move_and_clean(0.005)
    ctx.synchronize()
    t = ctx.phase_times()[0]
    best = t if best is None else {k: min(best[k], v) for k, v in t.items()}
out = ctx.get_particles()
import hashlib
print(os.environ.get("VVGPU_LIB", "default"), n, {k: round(v, 3) for k, v in best.items()}, "sum", round(sum(best.values()), 3),
      "hash", hashlib.blake2b(np.ascontiguousarray(out[:, :3]).tobytes(), digest_size=8).hexdigest())
