mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_near -c 6 -o gpurun_out/prof_near_all python tools/prof_one.py 1000000 1 > gpurun_out/ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_traverse|k_tree_swap|k_tree_bbox|k_sort_units" -c 12 -o gpurun_out/prof_tree python tools/prof_one.py 1000000 1 > gpurun_out/ncu_full3.log 2>&1
timeout 300 python -m pytest tests/test_dropin_cpp.py -x -q -m gpu > gpurun_out/pytest_dropin.log 2>&1
tail -3 gpurun_out/ncu_full2.log; tail -15 gpurun_out/pytest_dropin.log
