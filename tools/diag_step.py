"""Scratch: where does the resident bench loop lose time outside the phase timers?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from vvflow_b200 import capi, multigpu
n = 1_000_000
rec = bench.lamb_oseen_cloud(n)
ctx = capi.Context(0)
st = multigpu.ShardedStep(ctx, 0, 1, "cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
def loop(tag, do_flush, sampler, k=8):
    ctx.set_particles(rec)
    for _ in range(3): st.step(8, 0.0, bench.DBL_MAX, True, 1.0, 0.0, 0.005, 1000.0)
    ctx.phase_times()
    s = None
    if sampler:
        s = bench.ClockSampler(0); s.start()
    ts = []
    for _ in range(k):
        if do_flush: flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t_sub = []
        ctx.tree_build(8, 0.0, bench.DBL_MAX); t_sub.append(time.perf_counter())
        ctx.epsilon(True); t_sub.append(time.perf_counter())
        ctx.convective(1.0, 0.0, 0.005); t_sub.append(time.perf_counter())
        ctx.diffusive(1000.0, want_fric=False); t_sub.append(time.perf_counter())
        ctx.tree_destroy(); ctx.move_and_clean(0.005); t_sub.append(time.perf_counter())
        ms, la = ctx.phase_times()
        t1 = time.perf_counter()
        ts.append(((t1 - t0) * 1e3, sum(ms.values()), [round((b - a) * 1e3, 2) for a, b in zip([t0] + t_sub, t_sub + [t1])], {k_: round(v, 2) for k_, v in ms.items()}))
    if s: s.stop_flag = True; s.join()
    print(tag)
    for t in ts: print("   wall %.2f phases %.2f host-side %s %s" % t)
loop("plain", False, False)
loop("flush", True, False)
loop("flush+sampler", True, True)
