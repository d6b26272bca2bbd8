mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_traverse|k_tree_build_coop" -s 4 -c 4 -o gpurun_out/prof_trav -f python tools/prof_one.py 1000000 2 > gpurun_out/ncu_trav.log 2>&1
tail -2 gpurun_out/ncu_trav.log
