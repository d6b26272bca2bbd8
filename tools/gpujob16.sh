VVGPU_LIB=$PWD/vvflow_b200/lib/variants/libvvgpu_tt.so timeout 200 python tools/perf_probe.py 2>&1 | grep -E "tree timing|rep2" | tail -12
