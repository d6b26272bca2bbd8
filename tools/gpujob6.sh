mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_traverse" -c 4 -o gpurun_out/prof_trav python tools/prof_one.py 1000000 1 > gpurun_out/ncu_trav.log 2>&1
tail -3 gpurun_out/ncu_trav.log
