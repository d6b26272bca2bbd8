timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
VVGPU_LIB=$PWD/vvflow_b200/lib/variants/libvvgpu_tt.so timeout 200 python tools/perf_probe.py 2>&1 | grep -E "tree timing|rep2" | awk '!seen[$0]++' | tail -12
