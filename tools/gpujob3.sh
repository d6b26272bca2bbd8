mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/perf_probe.py > gpurun_out/probe.log 2>&1
tail -25 gpurun_out/pytest_gpu.log; cat gpurun_out/probe.log
