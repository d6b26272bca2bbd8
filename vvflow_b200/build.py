"""Builds vvflow_b200/lib/libvvgpu.so (sm_100a only) with nvcc. In-tree, so the .so travels to the GPU box."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "vvgpu.cu")
OUT = os.path.join(HERE, "lib", "libvvgpu.so")
OUT_TMA = os.path.join(HERE, "lib", "libvvgpu_tma.so")   # K4 with TMA-staged source tiles (a measured, slower variant)
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [
    os.path.join(HERE, "..", "include", "vvgpu.h")]


def nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvvgpu.so cannot be built (there is no CPU fallback)")


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False, variants=False):
    if not force and not stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    if variants:
        subprocess.check_call([c if c != OUT else OUT_TMA for c in cmd] + ["-DVV_CV_TMA=1"])
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
