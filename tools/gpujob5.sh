mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_near$" -s 2 -c 1 -o gpurun_out/prof_diff python tools/prof_one.py 1000000 1 > gpurun_out/ncu_diff.log 2>&1
tail -3 gpurun_out/ncu_diff.log
