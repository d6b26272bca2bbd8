#!/bin/bash
# final single-GPU validation of the round: tests, smoke, benches, ncu evidence
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest.txt; cat $O/pytest.txt | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
timeout 400 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; python tools/bench_line.py < $O/bench_1gpu.json
timeout 200 python bench.py --workload cyl --no-cpu-baseline > $O/bench_cyl.json 2> /dev/null; python tools/bench_line.py < $O/bench_cyl.json
timeout 200 python bench.py --particles 4000000 --steps 5 --no-cpu-baseline > $O/bench_4M.json 2> /dev/null; python tools/bench_line.py < $O/bench_4M.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python tools/prof_one.py 1000000 2 > $O/launches.log 2>&1
python tools/launch_summary.py $O/launches.csv 2 > $O/launch_summary.txt; head -12 $O/launch_summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_conv|k_diff|k_traverse_cta|k_near|k_tree_top|k_tree_sub|k_tree_topsweep|k_tree_relocate|k_heavy_pack' --launch-skip 20 -o $O/full python tools/prof_one.py 1000000 2 > $O/full.log 2>&1
ls -la $O
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/grp8.csv python tools/prof_group.py 1000000 8 2 > $O/grp8.log 2>&1
python tools/launch_by_stream.py $O/grp8.csv 2 > $O/grp8.txt; head -12 $O/grp8.txt
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/cyl_launches.csv python tools/prof_cyl.py 1000000 2 > $O/cyl_launches.log 2>&1
python tools/launch_summary.py $O/cyl_launches.csv 2 > $O/cyl_summary.txt; head -8 $O/cyl_summary.txt
