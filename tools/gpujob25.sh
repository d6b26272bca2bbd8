timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, cases
from vvflow_b200 import capi
ctx = capi.Context(0)
xyg = cases.cloud(1_000_000, "gauss", "equal", seed=12345)
ctx.set_particles_xyg(xyg); ctx.tree_build(8, 0.0); ctx.epsilon(True)
g = np.linspace(-3, 3, 1000)
pts = np.stack(np.meshgrid(g, g), -1).reshape(-1, 2)
for rep in range(3):
    t0 = time.perf_counter(); v = ctx.velocity_at(pts, 1.0, 0.0, 0.005); dt = time.perf_counter() - t0
    print(f"velocity_at: {pts.shape[0]} points on N=1M in {dt*1e3:.1f} ms = {pts.shape[0]/dt/1e6:.2f} M points/s (host in/out included), finite={np.isfinite(v).all()}")
PY
