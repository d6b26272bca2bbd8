timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","e2e","checksum_sum_g")})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_2gpu.err").read()[-2500:])
PY
