timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 --particles 16000000 --no-cpu-baseline > gpurun_out/bench_1gpu_16M.json 2> gpurun_out/bench_1gpu_16M.err; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_1gpu_16M.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","phase_ms","e2e")}, d["roofline"]["frac"], d["config"]["leaves"], d["config"]["near_pairs_per_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_1gpu_16M.err").read()[-1500:])
PY
