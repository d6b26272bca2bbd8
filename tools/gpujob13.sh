for v in cv2 cv3 df6 df12 eps6 eps12; do echo "== $v"; VVGPU_LIB=$PWD/vvflow_b200/lib/variants/libvvgpu_$v.so timeout 200 python tools/perf_probe.py 1000000 2>&1 | grep "rep2" | tail -1; done
