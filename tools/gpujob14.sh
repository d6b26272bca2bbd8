mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_diff$" -s 1 -c 1 -o gpurun_out/prof_diff -f python tools/prof_one.py 1000000 2 > gpurun_out/ncu_diff.log 2>&1
tail -2 gpurun_out/ncu_diff.log
