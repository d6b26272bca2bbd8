mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_traverse_cta$|^k_near$" -s 2 -c 2 -o gpurun_out/prof_trav2 -f python tools/prof_one.py 1000000 2 > gpurun_out/ncu_trav2.log 2>&1
tail -1 gpurun_out/ncu_trav2.log
