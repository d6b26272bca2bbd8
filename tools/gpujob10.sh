mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_near$" -s 3 -c 3 -o gpurun_out/prof_near_r1b -f python tools/prof_one.py 1000000 2 > gpurun_out/ncu_near.log 2>&1
tail -3 gpurun_out/ncu_near.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python tools/prof_one.py 1000000 2 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 1500 gpurun_out/bench_1gpu.json
