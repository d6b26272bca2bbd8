mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --particles 4000000 > gpurun_out/bench_2gpu_4M.json 2> gpurun_out/bench_2gpu_4M.err
timeout 600 python bench.py --steps 5 --warmup 3 --particles 4000000 --no-cpu-baseline > gpurun_out/bench_1gpu_4M.json 2> gpurun_out/bench_1gpu_4M.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_nocpu.json 2> gpurun_out/bench_1gpu_nocpu.err
for f in bench_1gpu_nocpu bench_2gpu bench_1gpu_4M bench_2gpu_4M; do echo == $f; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","phase_ms","e2e","gpu_launches")}, d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/$f.err").read()[-1500:])
PY
done
