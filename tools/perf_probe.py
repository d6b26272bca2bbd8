"""Scratch GPU probe: per-phase CUDA-event times of the hot path on synthetic clouds."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from vvflow_b200 import capi  # noqa: E402
if os.environ.get('VVGPU_LIB'):
    capi.LIB_PATH = os.environ['VVGPU_LIB']

ctx = capi.Context(0)
print("fp64 peak TFLOP/s:", ctx.fp64_peak())
for n, kind in [(100_000, "gauss"), (1_000_000, "gauss"), (4_000_000, "uniform")]:
    if len(sys.argv) > 1 and n > int(sys.argv[1]):
        continue
    xyg = cases.cloud(n, kind, "equal" if kind == "gauss" else "same", seed=12345)
    for rep in range(3):
        ctx.set_particles_xyg(xyg)
        t0 = time.time()
        ctx.tree_build(8, 0.0)
        ctx.epsilon(True)
        ctx.convective(1.0, 0.0, 0.005)
        ctx.diffusive(1000.0, want_fric=False)
        ctx.tree_destroy()
        ctx.move_and_clean(0.005)
        ctx.synchronize()
        wall = time.time() - t0
        ms, launches = ctx.phase_times()
        if rep == 0:
            ctx.set_particles_xyg(xyg)
            ctx.tree_build(8, 0.0)
            pairs, far = ctx.count_interactions()
            nn, nl, depth = ctx.tree_counts()
            ctx.tree_destroy()
            ctx.phase_times()
            print(f"N={n} nodes={nn} leaves={nl} depth={depth} near_pairs={pairs:.4g} far={far:.4g}")
        print(f"  rep{rep} wall={wall*1e3:.1f}ms launches={launches} " + " ".join(f"{k}={v:.2f}" for k, v in ms.items()),
              f"conv {pairs/ms['conv']/1e6:.1f} Gpairs/s")
