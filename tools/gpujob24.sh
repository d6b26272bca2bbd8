mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_conv$|^k_diff$|^k_near$|^k_tree_build_coop$|^k_traverse_cta$" -s 5 -c 6 -o gpurun_out/prof_r1f -f python tools/prof_one.py 1000000 2 > gpurun_out/ncu_r1e.log 2>&1
tail -1 gpurun_out/ncu_r1e.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python tools/prof_one.py 1000000 2 > gpurun_out/ncu_launch.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 1600 gpurun_out/bench_1gpu.json | cut -c1-1600
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 700 gpurun_out/bench_ref.json
