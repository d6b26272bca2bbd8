"""The C++ adapter (vvflow_b200/host/vvgpu_adapter.hpp) as a drop-in inside the reference's own step loop.

oracle/_ref/dropin_step is tests/host/dropin_step.cpp compiled against the reference's headers and
linked with the reference's objects (oracle/_ref/libvvref.so) and libvvgpu.so: the loop of
utils/vvflow/vvflow.cpp:198-266 for example/cyl_re600.lua where only the hot-path block (:246-257)
names vvgpu:: classes. It needs /root/reference to BUILD (done by `make -C oracle`), not to run.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_step")

# README.md:117-120 of the reference: time, force_hydro x / y / o (the only published known answers)
README_ROWS = [
    (0.00, "+3.140723e+01", None, None),
    (0.05, "+4.766549e-01", "+2.608255e-06", None),
    (0.10, "+8.190494e-01", "-1.868534e-03", "-5.548347e-05"),
    (0.15, "+7.309763e-01", "-1.069637e-04", "+5.027510e-05"),
]


def _need_bin():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/dropin_step not built (needs the reference headers; `make -C oracle`)")


def test_adapter_fails_loudly_without_gpu():
    """no CPU fallback: without a CUDA device the first hot-path call raises"""
    _need_bin()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = subprocess.run([BIN, "free", "1"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 2 and "no CPU fallback" in p.stdout, p.stdout + p.stderr


@pytest.mark.gpu
def test_free_run_reproduces_readme_rows():
    """cyl_re600 with the GPU hot path reproduces the reference's published force_hydro rows"""
    _need_bin()
    p = subprocess.run([BIN, "free", "4"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    rows = [l.split() for l in p.stdout.strip().splitlines()]
    assert len(rows) == 4
    for got, want in zip(rows, README_ROWS):
        assert abs(float(got[0]) - want[0]) < 1e-12
        for k in (1, 2, 3):
            if want[k] is not None:
                assert got[k] == want[k], (got, want)   # all 7 printed digits


@pytest.mark.gpu
def test_lockstep_against_reference_classes():
    """40 steps from identical state each step: order, g, _1_eps, merge and removal decisions
    bit-exact; positions, fric, gsum and dead-vortex sums within 1e-10 (checked inside the binary)"""
    _need_bin()
    p = subprocess.run([BIN, "lockstep", "40"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.strip().splitlines()[-1].startswith("OK"), p.stdout[-500:]


@pytest.mark.gpu
def test_points_and_slae_rhs_against_reference_classes():
    """SURVEY 8(f) rows 1 and 4 through the C++ adapter: after 12 lock-step steps of cyl_re600, velocity(p) on a
    60 x 40 raster and NodeInfluence() of all 350 segments against the reference's own classes in the same
    process (1e-10, checked inside the binary)"""
    _need_bin()
    p = subprocess.run([BIN, "points", "12"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    lines = p.stdout.strip().splitlines()
    assert lines[-1].startswith("OK") and lines[-2].startswith("points:"), p.stdout[-500:]


@pytest.mark.gpu
def test_resident_loop_without_cpu_particle_tree():
    """SURVEY 8(f) rows 1 + 2 through the C++ adapter: the vortex list lives on the device (shed vortices appended, no
    per-step upload or download), the SLAE right-hand side takes NodeInfluence from the device tree, so the reference's
    fill_matrix runs on a CPU tree of the body segments only. The README rows are reproduced, the free-running
    force_hydro stays within 1e-8 of an all-reference run beside it for 30 steps, and so does the final particle list."""
    _need_bin()
    p = subprocess.run([BIN, "resident", "30"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    lines = p.stdout.strip().splitlines()
    assert lines[-1].startswith("OK") and lines[-2].startswith("resident:"), p.stdout[-600:]
    rows = [l.split() for l in lines[:4]]
    for got, want in zip(rows, README_ROWS):
        for k in (1, 2, 3):
            if want[k] is not None:
                assert got[k] == want[k], (got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,steps", [("lockstep", 25), ("resident", 20)])
def test_two_ranks_from_one_process(mode, steps):
    """SURVEY 8(e) from inside the vvflow process: VVGPU_DEVICES lists one device per rank (here device 0 twice, so it
    runs on one GPU), the adapter drives every rank's context from its own host thread and the library's exchanges meet
    there. The same checks as with one rank: bit-exact order / merge / removal decisions, 1e-10 on every float."""
    _need_bin()
    env = dict(os.environ, VVGPU_DEVICES="0,0")
    p = subprocess.run([BIN, mode, str(steps)], capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    last = p.stdout.strip().splitlines()[-1]
    assert last.startswith("OK") and last.endswith("ranks=2"), p.stdout[-500:]
