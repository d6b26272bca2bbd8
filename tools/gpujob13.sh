for v in dff768 dff1024; do echo "== $v"; VVGPU_LIB=$PWD/vvflow_b200/lib/variants/libvvgpu_$v.so timeout 200 python tools/perf_probe.py 2>&1 | grep "rep2" | tail -2; done
