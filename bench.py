#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark (BASELINE.json): one full particle step
(tree build -> epsilon/merge -> convective Biot-Savart -> diffusive -> move/clean,
utils/vvflow/vvflow.cpp:246-257) on BASELINE config 2: a synthetic Lamb-Oseen vortex cloud, N = 1M.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--particles N]

Under torchrun (N > 1) one rank per GPU: targets are sharded, sources replicated by all-gathers.
Prints ONE JSON line on rank 0. `value` = steps/s with the particle state resident in HBM;
`e2e` = the same through the host boundary (48-byte TObj records in pinned host memory copied in
and out every step); `roofline` = the convective kernel against the measured FP64 pipe peak;
`cpu_baseline` = the reference's own code on this box's host cores on a bounded sample of leaves.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sim steps/s (full particle hot path; Biot-Savart interactions/s reported beside it)"
DBL_MAX = float(np.finfo(np.float64).max)
RE, DT, INF_VX, INF_VY, FAR = 1000.0, 0.005, 1.0, 0.0, 8


def lamb_oseen_cloud(n, seed=12345):
    """BASELINE config 2: x,y ~ N(0,1) (Lamb-Oseen blob, a = sqrt 2), g = 1/N, no bodies"""
    rng = np.random.default_rng(seed)
    rec = np.zeros((n, 6))
    rec[:, :2] = rng.standard_normal((n, 2))
    rec[:, 2] = 1.0 / n
    return rec


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md's clocks line), read
    in-process through NVML: spawning nvidia-smi per sample stalls the driver for milliseconds and
    showed up as +4.6 ms per step in the first runs of this bench."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.h, self.nv = index, [], False, None, None
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
            time.sleep(0.02)

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
                "how": "NVML in-process, 20 ms period, during both timed regions"}


# ------------------------------------------------------------------------------------ reference arm
def _omp_threads(n):
    """thread count of the OpenMP runtime the reference build is linked against"""
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


def reference_arm(args):
    """The reference's own CPU implementation (oracle/_ref: libvvhd sources compiled unmodified; else
    the C restatement) on this box's host cores. One step = tree build on the full cloud + epsilon,
    convective, diffusive on every `stride`-th leaf, extrapolated to all leaves + move_and_clean."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n
    rec = lamb_oseen_cloud(n)
    from oracle import pyref
    kind = "reference" if pyref.available() else "port"
    stride = max(1, args.ref_stride)
    # Threads: the reference recommends OMP_NUM_THREADS=1 (README.md:70, pytest/conftest.py:24) because its
    # OpenMP loops over leaves do not scale. Both settings are tried on a coarser sample and
    # the faster one is used for the measured steps, so the arm runs with all the threads it can USE.
    cores, calib = 1, {}
    if kind == "reference":
        ncpu = os.cpu_count() or 1
        for th in sorted({1, ncpu}):
            _omp_threads(th)
            r = pyref.Ref(re=RE, dt=DT, inf_vx=INF_VX, inf_vy=INF_VY)
            r.set_list(rec[:, :3])
            r.tree_params(FAR, 0.0, DBL_MAX)
            t0 = time.perf_counter(); r.tree_build(); tb = time.perf_counter() - t0
            r.sample_leaves(stride * 8, 0)
            t0 = time.perf_counter(); r.epsilon(True); r.convective(); r.diffusive(); tl = time.perf_counter() - t0
            r.tree_destroy(); r.close()
            calib[str(th)] = tb + 8 * stride * tl
        cores = int(min(calib, key=calib.get))
        _omp_threads(cores)
    times = []
    pairs_rate = []
    for it in range(args.warmup + args.steps):
        ph = it % stride
        if kind == "reference":
            r = pyref.Ref(re=RE, dt=DT, inf_vx=INF_VX, inf_vy=INF_VY)
            r.set_list(rec[:, :3])
            r.tree_params(FAR, 0.0, DBL_MAX)
            t0 = time.perf_counter(); r.tree_build(); t_build = time.perf_counter() - t0
            r.sample_leaves(stride, ph)
            npairs, _, _ = r.count_interactions()
            t0 = time.perf_counter(); r.epsilon(True); t_eps = time.perf_counter() - t0
            t0 = time.perf_counter(); r.convective(); t_conv = time.perf_counter() - t0
            t0 = time.perf_counter(); r.diffusive(); t_diff = time.perf_counter() - t0
            t0 = time.perf_counter(); r.tree_destroy(); r.move_and_clean(True); t_move = time.perf_counter() - t0
            r.close()
        else:
            from oracle import pyport
            p = pyport.Port(rec48=rec)
            t0 = time.perf_counter(); p.tree_build(FAR, 0.0, DBL_MAX); t_build = time.perf_counter() - t0
            nl = p.tree.contents.n_leaves
            npairs = sum(p.count_interactions(l, l + 1)[0] for l in range(ph, nl, stride))
            p.sample_leaves(stride, ph)
            t0 = time.perf_counter(); p.epsilon(True); t_eps = time.perf_counter() - t0
            t0 = time.perf_counter(); p.convective(INF_VX, INF_VY, DT); t_conv = time.perf_counter() - t0
            t0 = time.perf_counter(); p.diffusive(RE); t_diff = time.perf_counter() - t0
            p.sample_leaves(1, 0)
            t0 = time.perf_counter(); p.tree_destroy(); p.move_and_clean(DT); t_move = time.perf_counter() - t0
        step = t_build + stride * (t_eps + t_conv + t_diff) + t_move
        if it >= args.warmup:
            times.append(step)
            pairs_rate.append(npairs / t_conv)
    ms = 1e3 * float(np.mean(times))
    val = 1e3 / ms
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"Lamb-Oseen vortex cloud N={n}, one velocity+diffusion step (BASELINE configs[1])",
                   "n_particles": n, "re": RE, "dt": DT, "far_criteria": FAR},
        "interactions_per_s": float(np.mean(pairs_rate)),
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": kind,
                         "sample": f"full tree build + every {stride}-th leaf for epsilon/convective/diffusive, "
                                   f"times x{stride}; OpenMP threads = the faster of 1 (the reference's recommended "
                                   f"setting) and all {os.cpu_count()} host cores on a x8 coarser sample",
                         "thread_calibration_s_per_step": calib},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ our arm
def cpu_baseline_sample(n, rec, stride):
    from oracle import pyref
    if not pyref.available():
        return None
    os.environ["OMP_NUM_THREADS"] = "1"
    r = pyref.Ref(re=RE, dt=DT, inf_vx=INF_VX, inf_vy=INF_VY)
    r.set_list(rec[:, :3])
    r.tree_params(FAR, 0.0, DBL_MAX)
    t0 = time.perf_counter(); r.tree_build(); t_build = time.perf_counter() - t0
    r.sample_leaves(stride, 0)
    npairs, _, _ = r.count_interactions()
    t0 = time.perf_counter(); r.epsilon(True); t_eps = time.perf_counter() - t0
    t0 = time.perf_counter(); r.convective(); t_conv = time.perf_counter() - t0
    t0 = time.perf_counter(); r.diffusive(); t_diff = time.perf_counter() - t0
    t0 = time.perf_counter(); r.tree_destroy(); r.move_and_clean(True); t_move = time.perf_counter() - t0
    r.close()
    step = t_build + stride * (t_eps + t_conv + t_diff) + t_move
    return {"value": 1.0 / step, "unit": "steps/s", "cores": 1, "kind": "reference",
            "sample": f"full tree build + every {stride}-th leaf for epsilon/convective/diffusive, times x{stride} "
                      f"(OMP_NUM_THREADS=1, the reference's recommended setting)",
            "interactions_per_s": npairs / t_conv,
            "phase_s": {"build": t_build, "eps": stride * t_eps, "conv": stride * t_conv, "diff": stride * t_diff,
                        "move": t_move}}


def ours(args):
    import torch
    import torch.distributed as dist
    from vvflow_b200 import capi, multigpu

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))

    n = args.n
    rec = lamb_oseen_cloud(n)
    ctx = capi.Context(local)
    stepper = multigpu.ShardedStep(ctx, rank, world, device)
    host_in = torch.empty((n, 6), dtype=torch.float64).pin_memory()
    host_out = torch.empty((n, 6), dtype=torch.float64).pin_memory()
    host_in.numpy()[:] = rec
    flush = torch.empty(32 << 20, dtype=torch.float64, device=device)  # 256 MB > 126 MB L2

    def barrier():
        ctx.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        return stepper.step(FAR, 0.0, DBL_MAX, True, INF_VX, INF_VY, DT, RE)

    def flush_l2():
        flush.zero_()

    # ---- resident: the state stays in HBM from step to step (the simulation loop itself)
    ctx.set_particles(rec)
    for _ in range(args.warmup):
        flush_l2()   # also warms the fill kernel up: its first launch costs ~25 ms (lazy module load)
        one_step()
    ctx.phase_times()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    phase_sum = {k: 0.0 for k in capi.PHASES}
    launches = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    dbg = os.environ.get("VV_BENCH_DEBUG")
    for _ in range(args.steps):
        ta = time.perf_counter()
        flush_l2()
        torch.cuda.synchronize()
        tb = time.perf_counter()
        one_step()
        tc = time.perf_counter()
        ms, la = ctx.phase_times()  # synchronises the library's stream
        if dbg:
            print(f"[dbg] flush {1e3*(tb-ta):.2f} step {1e3*(tc-tb):.2f} sync {1e3*(time.perf_counter()-tc):.2f} phases {sum(ms.values()):.2f}", file=sys.stderr)
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

    if dbg:
        print(f"[dbg] rank {rank}: " + " ".join(f"{k}={v / args.steps:.2f}" for k, v in phase_sum.items())
              + f" sum={sum(phase_sum.values()) / args.steps:.2f} wall/step={wall_ms / args.steps:.2f}", file=sys.stderr)
    # interaction counts of the final state's tree (work done per step)
    ctx.tree_build(FAR, 0.0, DBL_MAX)
    near_pairs, far_nodes = ctx.count_interactions()   # totals over the ranks (summed inside the library)
    local_pairs = near_pairs / world                    # the groups are dealt round-robin: every rank has ~1/world of them
    nn, nl, depth = ctx.tree_counts()
    ctx.tree_destroy()
    ctx.phase_times()
    conv_ms_max = phase_sum["conv"] / args.steps
    if world > 1:
        t = torch.tensor([conv_ms_max], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        conv_ms_max = float(t[0].item())

    # ---- e2e: host records in, host records out, every step. On one GPU the whole list crosses PCIe both ways.
    # On N GPUs every rank moves ONE SLICE of the records each way (its own pinned buffers) and the slices are
    # all-gathered over NVLink into a device buffer that is handed to the library; pushing the full list through
    # every rank's PCIe link cost 6 ms per step at 8 GPUs.
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    rec_bytes = 48

    def e2e_in():
        if world == 1:
            ctx.set_particles_ptr(host_in.data_ptr(), n)
        else:   # one slice per rank over PCIe, gathered over NVLink inside the library
            ctx.set_particles_slice_ptr(host_in.data_ptr() + lo * rec_bytes, lo, hi - lo, n)

    def e2e_out():
        if world == 1:
            return ctx.get_particles_ptr(host_out.data_ptr(), n), n
        k = ctx.n                                                   # every rank holds the whole (replicated) result
        a, b = (k * rank) // world, (k * (rank + 1)) // world       # ... and brings its share back to the host
        ctx.get_particles_range_ptr(host_out.data_ptr() + a * rec_bytes, a, b - a)
        return k, b - a

    for _ in range(min(args.warmup, 2)):
        e2e_in()
        one_step()
        e2e_out()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush_l2()
        torch.cuda.synchronize()
        e2e_in()
        one_step()
        nout, nback = e2e_out()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(t[0].item()) / args.steps
    sampler.stop_flag = True
    sampler.join(timeout=3)
    if world == 1:
        checksum = float(host_out.numpy()[:nout, 2].sum())
    else:   # every rank's share of sum(g), added up
        a, b = (nout * rank) // world, (nout * (rank + 1)) // world
        t = torch.tensor([float(host_out.numpy()[a:b, 2].sum())], dtype=torch.float64, device=device)
        dist.all_reduce(t)
        checksum = float(t[0].item())

    # ---- roofline of the dominant kernel family: K4 convective, FP64 pipe
    fp64_peak = ctx.fp64_peak()
    conv_ms = phase_sum["conv"] / args.steps
    # targets are sharded: rank 0 reports its own kernel on its own pairs
    flops = 11.0 * local_pairs
    achieved = flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_sample(n, rec, args.ref_stride)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": 1e3 / ms_per_step, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Lamb-Oseen vortex cloud N={n}, one velocity+diffusion step (BASELINE configs[1])",
                       "n_particles": n, "re": RE, "dt": DT, "far_criteria": FAR, "leaves": nl, "tree_depth": depth,
                       "near_pairs_per_step": near_pairs, "far_nodes_per_step": far_nodes,
                       "parallelism": f"target-sharded x{world}, sources replicated by all-gather" if world > 1 else "1 GPU",
                       "l2": "256 MB memset between steps (inside the timed region) flushes the 126 MB L2"},
            "interactions_per_s": near_pairs / (conv_ms_max * 1e-3) if conv_ms_max > 0 else None,
            "phase_ms": {k: v / args.steps for k, v in phase_sum.items()},
            "wall_ms_per_step": wall_total_ms / args.steps,
            "e2e": {"value": 1e3 / e2e_ms_per_step, "unit": "steps/s", "ms_per_step": e2e_ms_per_step,
                    "h2d_bytes_per_step": 48 * n, "d2h_bytes_per_step": 48 * int(nout),
                    "io": "one GPU: the whole list each way" if world == 1 else
                          f"each of the {world} ranks moves 1/{world} of the records each way (bytes are the totals over "
                          "ranks); slices all-gathered over NVLink"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "kernel": "k_conv (K4 convective near field)",
                         "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak else None,
                         # ncu dram__bytes_read + dram__bytes_write of this kernel at N=1M, per launch
                         # (profiles/r1_ncu_kernels_v7.txt: 97.2 MB + 13.2 MB)
                         "traffic": 1.10e8 if n == 1_000_000 and world == 1 else None,
                         "ncu_fp64_pipe_pct": 58.9 if n == 1_000_000 and world == 1 else None,
                         "note": "FP64-pipe bound (no tensor cores: N-body gather). achieved = 11 flop x near pairs / "
                                 "CUDA-event time of the phase on the library's stream; peak = DFMA micro-benchmark "
                                 "measured in this run (MEASURED_PEAKS.json holds HBM %s GB/s and bf16 only). "
                                 "HBM traffic of the kernel is ~0.11 GB vs 39.7 GFLOP: compute-bound. ncu_fp64_pipe_pct = "
                                 "sm__inst_executed_pipe_fp64 of the committed ncu capture (not measured in this run): the "
                                 "pipe is busier than frac says because a pair costs 9 FP64 instructions, of which only "
                                 "7 are FMAs, for the 11 flop the metric counts." % peaks.get("hbm_gbs")},
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "checksum_sum_g": checksum,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", dest="n", type=int, default=1_000_000,
                    help="N of the synthetic cloud (not `--n`: torchrun would read that as one of its own options)")
    ap.add_argument("--ref-stride", type=int, default=64, help="CPU arms: every stride-th leaf is evaluated")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
