for v in t512b8 t512b4 t256b8; do echo "== $v"; VVGPU_LIB=$PWD/vvflow_b200/lib/variants/libvvgpu_$v.so timeout 200 python tools/perf_probe.py 2>&1 | grep -E "rep2|rror" | tail -3; done
