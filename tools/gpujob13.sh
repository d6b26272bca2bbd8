timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/perf_probe.py 2>&1 | grep "rep2"
for v in eps16 eps8m4; do echo "== $v"; VVGPU_LIB=$PWD/vvflow_b200/lib/variants/libvvgpu_$v.so timeout 200 python tools/perf_probe.py 2>&1 | grep "rep2"; done
