#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark (BASELINE.json): one full particle step
(tree build -> epsilon/merge -> convective Biot-Savart -> diffusive -> move/clean,
utils/vvflow/vvflow.cpp:246-257).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload lamb|uniform|cyl] [--particles N]

Default workload = BASELINE configs[1]: a synthetic Lamb-Oseen vortex cloud, N = 1M. `uniform` is configs[4]
(x, y ~ U[0,1)^2, g ~ U[0.5,1]/N); `cyl` is a body case: the bundled cylinder (350 segments, tree parameters of
vvflow.cpp:200-203) inside ~N mixed-sign particles, so that merging, the wall passes, segment diffusion + fric and
in-body removal all run.

Under torchrun (N > 1) one rank per GPU: targets are sharded inside the library, sources replicated by its own
all-gathers. Prints ONE JSON line on rank 0. `value` = steps/s with the particle state resident in HBM;
`e2e` = the same through the host boundary (48-byte TObj records in pinned host memory copied in and out every
step); `roofline` = the convective kernel against the measured FP64 pipe peak; `cpu_baseline` = the reference's own
code on this box's host cores on a bounded sample of leaves. `--impl reference` times the reference's own code.
"""
import argparse
import glob
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sim steps/s (full particle hot path; Biot-Savart interactions/s reported beside it)"
DBL_MAX = float(np.finfo(np.float64).max)
FAR = 8


# ------------------------------------------------------------------------------------ workloads
def cylinder_points(R=0.5, nseg=350):
    """gen_cylinder = gen_arc_N(c, R, 2pi -> 0, N), utils/vvflow/gen_cylinder.cpp:33-41, gen_body.cpp:74-88"""
    i = np.arange(nseg, dtype=np.float64)
    a = 2 * np.pi + (0 - 2 * np.pi) * i / nseg
    return np.stack([R * np.cos(a), R * np.sin(a)], axis=1)


def cyl_wake_cloud(n, dl=2 * np.pi * 0.5 / 350, seed=777):
    """A wake-like cloud with the spacing a running simulation has: CalcEpsilonFast merges neighbours closer than
    0.4 dl sqrt(1 + dist) (MEpsilonFast.cpp:33), so the particle spacing of a real wake grows like sqrt(1 + dist) and a
    Poisson cloud of that density would merge (and fill the 5 dl leaves with) most of its particles. Here: a jittered
    lattice in coordinates stretched by s(x) = a sqrt(x + 0.5), a = 1.5 x the merge radius at the wall, in a wake of
    half-width 0.3 + 0.15 x, + 20 rings around the body + a few particles inside it. Upper half negative, lower
    positive, 3 % of the signs flipped (those merge by the sign rule, :164-166)."""
    rng = np.random.default_rng(seed + n)
    a = 1.5 * 0.4 * dl

    def wake(X):
        u = np.arange((2 / a) * np.sqrt(1.1), (2 / a) * np.sqrt(X + 0.5), 1.0)
        x = (a * u / 2) ** 2 - 0.5
        sx = a * np.sqrt(x + 0.5)
        half = np.floor((0.3 + 0.15 * x) / sx).astype(np.int64)
        cnt = 2 * half + 1
        col = np.repeat(np.arange(u.shape[0]), cnt)
        v = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt) - half[col]
        return x[col], v * sx[col], sx[col]

    rings = np.concatenate([np.stack([r * np.cos(t), r * np.sin(t)], 1) for r in 0.5 + 0.006 + a * np.arange(20)
                            for t in [np.arange(0, 2 * np.pi, a / r)]])
    inside = rng.uniform(-0.3, 0.3, (150, 2))
    need = n - rings.shape[0] - inside.shape[0]
    lo, hi = 1.0, 4000.0
    for _ in range(60):      # wake length that yields ~1.02 x the particles still needed
        mid = 0.5 * (lo + hi)
        if wake(mid)[0].shape[0] < 1.02 * need: lo = mid
        else: hi = mid
    x, y, sx = wake(hi)
    keep = rng.permutation(x.shape[0])[:need]
    keep.sort()
    x, y, sx = x[keep], y[keep], sx[keep]
    x = x + rng.uniform(-0.25, 0.25, x.shape[0]) * sx
    y = y + rng.uniform(-0.25, 0.25, x.shape[0]) * sx
    xy = np.concatenate([rings, inside, np.stack([x, y], 1)])[:n]
    rec = np.zeros((n, 6))
    rec[: xy.shape[0], :2] = xy
    sign = -np.sign(rec[:, 1]) + (rec[:, 1] == 0)
    flip = rng.uniform(0, 1, n) < 0.03
    rec[:, 2] = np.where(flip, -sign, sign) * rng.uniform(0.5, 1.0, n) * (10.0 / n)
    order = rng.permutation(n)     # a running simulation's list is not spatially sorted either
    return rec[order]


def make_workload(name, n):
    """inputs of one step: records (n, 6), body corner lists, physical parameters. numpy default_rng: both arms of
    the bench read the same bits (SURVEY 8(d) names an mt19937_64 file; the generator does not matter for parity
    as long as both sides consume identical records, which they do here)."""
    w = {"name": name, "n": n, "bodies": [], "inf": (1.0, 0.0), "merge": True}
    rec = np.zeros((n, 6))
    if name == "lamb":       # BASELINE configs[1]: x,y ~ N(0,1) (Lamb-Oseen blob, a = sqrt 2), g = 1/N, no bodies
        rng = np.random.default_rng(12345)
        rec[:, :2] = rng.standard_normal((n, 2))
        rec[:, 2] = 1.0 / n
        w.update(re=1000.0, dt=0.005, label=f"Lamb-Oseen vortex cloud N={n}, one velocity+diffusion step (BASELINE configs[1])")
    elif name == "uniform":  # BASELINE configs[4], 5a: x,y ~ U[0,1)^2, g ~ U[0.5,1]/N, seed 12345+N
        rng = np.random.default_rng(12345 + n)
        rec[:, :2] = rng.uniform(0, 1, (n, 2))
        rec[:, 2] = rng.uniform(0.5, 1.0, n) / n
        w.update(re=1000.0, dt=0.005, label=f"uniform random vortex cloud N={n} (BASELINE configs[4], 5a)")
    elif name == "cyl":      # body case: cylinder R = 0.5, 350 segments (example/cyl_re600.lua) in a synthetic wake
        rec = cyl_wake_cloud(n)
        w["bodies"] = [cylinder_points(0.5, 350)]
        w.update(re=600.0, dt=0.05, label=f"cylinder R=0.5 (350 segments, example/cyl_re600.lua geometry) in a synthetic "
                                          f"wake of N={n} mixed-sign particles: merges, wall passes, fric, in-body removal")
    else:
        raise SystemExit(f"unknown workload {name}")
    w["rec"] = rec
    return w


def make_config(args, w):
    """the SAME dictionary on both arms (the driver compares them)"""
    return {"workload": w["label"], "n_particles": w["n"], "re": w["re"], "dt": w["dt"], "far_criteria": FAR,
            "inf_speed": list(w["inf"]), "n_segments": int(sum(b.shape[0] for b in w["bodies"])), "merge": True,
            "generator": "numpy default_rng, fixed seed; identical records on both arms"}


def state_hash(rec):
    """64-bit hash of the (x, y, g) bits of a particle list"""
    a = np.ascontiguousarray(rec[:, :3])
    return hashlib.blake2b(a.tobytes(), digest_size=8).hexdigest()


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md's clocks line), read
    in-process through NVML: spawning nvidia-smi per sample stalls the driver for milliseconds and
    showed up as +4.6 ms per step in the first runs of this bench."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.h, self.nv = index, [], False, None, None
        self.period = 0.02 if int(os.environ.get("WORLD_SIZE", "1")) == 1 else 0.1
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; NVML indexes the physical devices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def run(self):
        nv = self.nv
        while not self.stop_flag and self.h is not None:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.rows.append((sm, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        nv = self.nv
        sm = sorted(r[0] for r in self.rows)
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = [k for k, b in bits.items() if any(r[1] & b for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(self.rows),
                "how": "NVML in-process on rank 0, %d ms period, during both timed regions" % int(self.period * 1e3)}


# ------------------------------------------------------------------------------------ reference arm
def _omp_threads(n):
    """thread count of the OpenMP runtime the reference build is linked against"""
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


class RefStep:
    """one hot-path step of the reference's own code (oracle/_ref: libvvhd compiled unmodified) or, where that did not
    travel, of the C restatement; stride > 1 evaluates every stride-th leaf only (a bounded SAMPLE of the step)"""

    def __init__(self, w):
        from oracle import pyref
        self.w, self.pyref = w, pyref
        self.kind = "reference" if pyref.available() else "port"

    def run(self, stride=1, phase=0):
        w = self.w
        re, dt, (ivx, ivy) = w["re"], w["dt"], w["inf"]
        t = {}
        if self.kind == "reference":
            r = self.pyref.Ref(re=re, dt=dt, inf_vx=ivx, inf_vy=ivy)
            for b in w["bodies"]:
                r.add_polygon(b)
            r.set_list(w["rec"][:, :3])
            if w["bodies"]:
                r.tree_params(FAR)           # vvflow.cpp:200-203 from the average segment length
            else:
                r.tree_params(FAR, 0.0, DBL_MAX)
            t0 = time.perf_counter(); r.tree_build(); t["build"] = time.perf_counter() - t0
            if stride > 1:
                r.sample_leaves(stride, phase)
            npairs = r.count_interactions()[0]
            t0 = time.perf_counter(); r.epsilon(True); t["eps"] = time.perf_counter() - t0
            t0 = time.perf_counter(); r.convective(); t["conv"] = time.perf_counter() - t0
            t0 = time.perf_counter(); r.diffusive(); t["diff"] = time.perf_counter() - t0
            t0 = time.perf_counter(); r.tree_destroy(); r.move_and_clean(True); t["move"] = time.perf_counter() - t0
            r.close()
        else:
            from oracle import pyport
            if w["bodies"]:
                raise SystemExit("the reference build (oracle/_ref) did not travel: the port arm has no body generator")
            p = pyport.Port(rec48=w["rec"])
            t0 = time.perf_counter(); p.tree_build(FAR, 0.0, DBL_MAX); t["build"] = time.perf_counter() - t0
            nl = p.tree.contents.n_leaves
            npairs = sum(p.count_interactions(l, l + 1)[0] for l in range(phase, nl, stride))
            p.sample_leaves(stride, phase)
            t0 = time.perf_counter(); p.epsilon(True); t["eps"] = time.perf_counter() - t0
            t0 = time.perf_counter(); p.convective(ivx, ivy, dt); t["conv"] = time.perf_counter() - t0
            t0 = time.perf_counter(); p.diffusive(re); t["diff"] = time.perf_counter() - t0
            p.sample_leaves(1, 0)
            t0 = time.perf_counter(); p.tree_destroy(); p.move_and_clean(dt); t["move"] = time.perf_counter() - t0
        est = t["build"] + stride * (t["eps"] + t["conv"] + t["diff"]) + t["move"]
        wall = sum(t.values())
        return {"estimate_s": est, "wall_s": wall, "pairs_per_s": stride * npairs / (stride * t["conv"]) if t["conv"] > 0 else None,
                "phase_s": {k: (v * stride if k in ("eps", "conv", "diff") else v) for k, v in t.items()}}


def reference_arm(args):
    """The reference's own CPU implementation on this box's host cores. The timed steps are FULL, un-sampled steps
    (every leaf): as many of the K requested as fit the time budget, at least one; the sampled estimate of the same
    step (every 64th leaf x 64, what `cpu_baseline` of the other arm uses) is reported beside it with its ratio."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload, args.n)
    rs = RefStep(w)
    stride = max(1, args.ref_stride)
    # Threads: the reference recommends OMP_NUM_THREADS=1 (README.md:70, pytest/conftest.py:24) because its
    # OpenMP loops over leaves do not scale. Both settings are tried on a coarse sample and
    # the faster one is used for the measured steps, so the arm runs with all the threads it can USE.
    cores, calib = 1, {}
    if rs.kind == "reference":
        for th in sorted({1, os.cpu_count() or 1}):
            _omp_threads(th)
            calib[str(th)] = rs.run(stride * 8, 0)["estimate_s"]
        cores = int(min(calib, key=calib.get))
        _omp_threads(cores)
    sampled = [rs.run(stride, it % stride) for it in range(max(1, min(args.warmup, 3)))]   # also warms the allocator up
    t_begin = time.perf_counter()
    full = []
    while len(full) < max(1, args.steps):
        full.append(rs.run(1, 0))
        mean = float(np.mean([f["wall_s"] for f in full]))
        if time.perf_counter() - t_begin + mean > args.ref_budget_s:
            break
    full_s = float(np.mean([f["wall_s"] for f in full]))
    samp_s = float(np.mean([s["estimate_s"] for s in sampled]))
    val = 1.0 / full_s
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": len(full), "warmup": 0, "ms_per_step": 1e3 * full_s, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args, w),
        "steps_requested": args.steps,
        "full_step_s": full_s, "full_steps_timed": len(full),
        "sampled_estimate_s": samp_s, "sampled_over_full": samp_s / full_s,
        "phase_s": {k: float(np.mean([f["phase_s"][k] for f in full])) for k in full[0]["phase_s"]},
        "interactions_per_s": float(np.mean([f["pairs_per_s"] for f in full])),
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": rs.kind,
                         "sample": f"{len(full)} FULL un-sampled step(s) (every leaf; as many of the {args.steps} requested as "
                                   f"fit {args.ref_budget_s:.0f} s), OpenMP threads = the faster of 1 (the reference's "
                                   f"recommended setting) and all {os.cpu_count()} host cores on a coarse sample; no warm-up "
                                   f"step is timed or needed (CPU code), {len(sampled)} sampled step(s) ran before",
                         "thread_calibration_s_per_step": calib},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ our arm
def cpu_baseline_sample(w, stride):
    """the reference on ONE host thread on a bounded sample (every stride-th leaf), 10-30 s of CPU work"""
    rs = RefStep(w)
    if rs.kind != "reference":
        return None
    _omp_threads(1)
    s = rs.run(stride, 0)
    return {"value": 1.0 / s["estimate_s"], "unit": "steps/s", "cores": 1, "kind": "reference",
            "sample": f"full tree build + every {stride}-th leaf for epsilon/convective/diffusive, times x{stride} "
                      f"(OMP_NUM_THREADS=1, the reference's recommended setting); the reference arm "
                      f"(--impl reference) times full un-sampled steps",
            "interactions_per_s": s["pairs_per_s"], "phase_s": s["phase_s"]}


def ncu_record(n, world):
    """ncu numbers of K4 for this N from the newest committed capture (profiles/r*_ncu_k_conv_N*.json carries the
    commit it was taken at); not measured in this run"""
    if world != 1:
        return None
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_k_conv_N*.json"))):
        try:
            d = json.load(open(f))
        except Exception:
            continue
        if d.get("n_particles") == n:
            best = dict(d, file=os.path.relpath(f, ROOT))
    return best


def ours(args):
    import torch
    import torch.distributed as dist
    from vvflow_b200 import capi, multigpu, vvhd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries ONE JSON line: whatever libraries print there meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))

    w = make_workload(args.workload, args.n)
    n, rec = w["n"], w["rec"]
    re, dt, (ivx, ivy) = w["re"], w["dt"], w["inf"]
    ctx = capi.Context(local)
    multigpu.init_comm(ctx, rank, world)       # the data plane is the library's; torch only carries the NCCL id
    S = vvhd.Space(ctx=ctx)
    S.BodyList = [vvhd.TBody(b) for b in w["bodies"]]
    ctx.set_bodies(*S._pack_bodies())
    dl = S.average_segment_length()
    tmin, tmax = (dl * 5, dl * 100) if dl > 0 else (0.0, DBL_MAX)     # vvflow.cpp:200-203
    reseed = bool(w["bodies"])    # merges / removals change N: every step starts from the same records again
    want_fric = bool(w["bodies"])

    host_in = torch.empty((n, 6), dtype=torch.float64).pin_memory()
    host_out = torch.empty((n, 6), dtype=torch.float64).pin_memory()
    host_in.numpy()[:] = rec
    dev_seed = torch.from_numpy(rec).to(device) if reseed else None
    flush = torch.empty(32 << 20, dtype=torch.float64, device=device)  # 256 MB > 126 MB L2
    stats = {"merged": 0, "cleaned": 0, "merge_rounds": 0}

    def barrier():
        ctx.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        if reseed:
            ctx.set_particles_ptr(dev_seed.data_ptr(), n)      # device-to-device
        ctx.tree_build(FAR, tmin, tmax)
        stats["merged"] = ctx.epsilon(True)
        stats["merge_rounds"] = ctx.merge_rounds()
        ctx.convective(ivx, ivy, dt)
        ctx.diffusive(re, want_fric=want_fric)
        ctx.tree_destroy()
        stats["cleaned"] = ctx.move_and_clean(dt)["cleaned"]

    def flush_l2():
        flush.zero_()

    # ---- resident: the state stays in HBM from step to step (the simulation loop itself)
    ctx.set_particles(rec)
    for _ in range(args.warmup):
        flush_l2()   # also warms the fill kernel up: its first launch costs ~25 ms (lazy module load)
        one_step()
    ctx.phase_times()
    ctx.host_syncs()
    sampler = ClockSampler(local)
    if rank != 0 or os.environ.get("VV_BENCH_NOSAMPLER"):
        sampler.h = None     # NVML queries from 8 processes stalled rank 0's launches (+4 ms per step): rank 0 samples alone
    sampler.start()
    barrier()
    phase_sum = {k: 0.0 for k in capi.PHASES}
    launches = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        flush_l2()
        torch.cuda.synchronize()
        one_step()
        syncs = ctx.host_syncs()    # host waits inside the step's calls (read-backs), counted by the library
        ms, la = ctx.phase_times()  # synchronises the library's stream
        for k in phase_sum:
            phase_sum[k] += ms[k]
        launches += la
    ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([max(dev_ms, 0.0), wall_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # CUDA events, max over ranks. The events sit on torch's stream while the library runs on its own, but every
    # step ends with a host synchronisation of the library's stream (phase_times), so ev1 is recorded after all of
    # the step's work; the wall clock around the same region (barrier + synchronize on both sides) is kept beside it.
    total_ms = float(t[0].item())
    wall_total_ms = float(t[1].item())
    ms_per_step = total_ms / args.steps
    phase_ms = {k: v / args.steps for k, v in phase_sum.items()}
    if world > 1:   # per phase the slowest rank
        t = torch.tensor([phase_ms[k] for k in capi.PHASES], dtype=torch.float64, device=device)
        tmax_ = t.clone(); dist.all_reduce(tmax_, op=dist.ReduceOp.MAX)
        tmin_ = t.clone(); dist.all_reduce(tmin_, op=dist.ReduceOp.MIN)
        phase_ms_max = dict(zip(capi.PHASES, tmax_.tolist()))
        phase_ms_min = dict(zip(capi.PHASES, tmin_.tolist()))
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        phase_by_rank = [[round(v, 3) for v in a.tolist()] for a in allr]
    else:
        phase_ms_max = phase_ms_min = phase_ms
        phase_by_rank = None

    # interaction counts of one step's tree (work done per step)
    if reseed:
        ctx.set_particles_ptr(dev_seed.data_ptr(), n)
    ctx.tree_build(FAR, tmin, tmax)
    near_pairs, far_nodes = ctx.count_interactions()   # totals over the ranks (summed inside the library)
    nn, nl, depth = ctx.tree_counts()
    ctx.tree_destroy()
    ctx.phase_times()

    # ---- e2e: host records in, host records out, every step. On one GPU the whole list crosses PCIe both ways.
    # On N GPUs every rank moves ONE SLICE of the records each way (its own pinned buffers); the library gathers the
    # uploaded slices over NVLink (vvgpu_set_particles_slice).
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world

    def e2e_in():
        if world == 1:
            ctx.set_particles_ptr(host_in.data_ptr(), n)
        else:
            ctx.set_particles_slice_ptr(host_in.data_ptr() + lo * 48, lo, hi - lo, n)

    def e2e_out():
        k = ctx.n                                                   # every rank holds the whole (replicated) result
        a, b = (k * rank) // world, (k * (rank + 1)) // world       # ... and brings its share back to the host
        ctx.get_particles_range_ptr(host_out.data_ptr() + a * 48, a, b - a)
        return k, a, b

    reseed_saved, reseed = reseed, False     # the host upload IS the re-seed
    for _ in range(min(args.warmup, 2)):
        e2e_in(); one_step(); e2e_out()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush_l2()
        torch.cuda.synchronize()
        e2e_in()
        one_step()
        nout, oa, ob = e2e_out()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    reseed = reseed_saved
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(t[0].item()) / args.steps
    sampler.stop_flag = True
    sampler.join(timeout=3)
    # the state after ONE step from the input records (what every e2e iteration computes): identical on every rank and
    # for every number of GPUs
    final = ctx.get_particles()
    h = state_hash(final)
    hashes_equal = True
    if world > 1:
        hv = torch.tensor([int(h, 16) >> 1], dtype=torch.int64, device=device)
        hs = [torch.zeros_like(hv) for _ in range(world)]
        dist.all_gather(hs, hv)
        hashes_equal = all(int(x.item()) == int(hv.item()) for x in hs)
    checksum = float(final[:, 2].sum())

    # ---- roofline of the dominant kernel family: K4 convective, FP64 pipe
    fp64_peak = ctx.fp64_peak()
    conv_ms = phase_ms["conv"]
    local_pairs = near_pairs / world            # groups are dealt round-robin: every rank has ~1/world of the pairs
    flops = 11.0 * local_pairs
    achieved = flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    clocks = sampler.summary()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_sample(w, args.ref_stride)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        ncu = ncu_record(n, world) if args.workload == "lamb" else None
        line = {
            "metric": METRIC, "value": 1e3 / ms_per_step, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args, w),
            "workload_stats": {"leaves": nl, "tree_depth": depth, "nodes": nn, "near_pairs_per_step": near_pairs,
                               "far_nodes_per_step": far_nodes, "merged_last_step": stats["merged"],
                               "merge_rounds_last_step": stats["merge_rounds"], "cleaned_last_step": stats["cleaned"],
                               "host_readbacks_last_step": int(syncs),
                               "parallelism": (f"targets sharded over {world} ranks inside libvvgpu (leaf groups dealt block-"
                                               f"cyclically), sources replicated by its ncclAllGather exchanges") if world > 1 else "1 GPU",
                               "l2": "256 MB memset between steps (inside the timed region) flushes the 126 MB L2",
                               "reseed": "every step restarts from the input records (device-to-device copy inside the timed "
                                         "region)" if reseed else "the state evolves from step to step"},
            "interactions_per_s": near_pairs / (phase_ms_max["conv"] * 1e-3) if phase_ms_max["conv"] > 0 else None,
            "phase_ms": phase_ms, "phase_ms_max_over_ranks": phase_ms_max, "phase_ms_min_over_ranks": phase_ms_min,
            "phase_ms_by_rank": phase_by_rank,
            "wall_ms_per_step": wall_total_ms / args.steps,
            "e2e": {"value": 1e3 / e2e_ms_per_step, "unit": "steps/s", "ms_per_step": e2e_ms_per_step,
                    "h2d_bytes_per_step": 48 * n, "d2h_bytes_per_step": 48 * int(nout),
                    "io": "one GPU: the whole list each way" if world == 1 else
                          f"each of the {world} ranks moves 1/{world} of the records each way (bytes are the totals over "
                          "ranks); the uploaded slices are gathered over NVLink inside the library"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "kernel": "k_conv (K4 convective near field)",
                         "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak else None,
                         "peak_how": "DFMA micro-benchmark (vvgpu_fp64_peak) run in this process right after the timed "
                                     "regions; MEASURED_PEAKS.json holds HBM %s GB/s and bf16 only" % peaks.get("hbm_gbs"),
                         "peak_sm_mhz": clocks.get("sm_mhz"),
                         "peak_nominal": 148 * 64 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12,
                         "traffic": (ncu["dram_bytes_read"] + ncu["dram_bytes_write"]) if ncu else None,
                         "ncu_fp64_pipe_pct": ncu["fp64_pipe_pct"] if ncu else None,
                         "ncu_capture": {k: ncu[k] for k in ("file", "commit", "duration_us")} if ncu else None,
                         "note": "FP64-pipe bound (no tensor cores: N-body gather). achieved = 11 flop x near pairs of this "
                                 "rank / CUDA-event time of the phase on the library's stream. traffic / ncu_fp64_pipe_pct "
                                 "come from the committed ncu capture named in ncu_capture (not measured in this run). The "
                                 "pipe is busier than frac says: a pair costs 9 FP64 instructions, 7 of them FMAs, for the "
                                 "11 flop the metric counts."},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "state_hash": h, "state_hash_same_on_all_ranks": hashes_equal,
            "state_hash_of": "(x, y, g) bits after ONE step from the input records; the same for every number of GPUs",
            "checksum_sum_g": checksum,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lamb", choices=["lamb", "uniform", "cyl"])
    ap.add_argument("--particles", dest="n", type=int, default=1_000_000,
                    help="N of the synthetic cloud (not `--n`: torchrun would read that as one of its own options)")
    ap.add_argument("--ref-stride", type=int, default=64, help="CPU sample: every stride-th leaf is evaluated")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="reference arm: time budget for the full steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
