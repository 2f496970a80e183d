python bench.py --kernel-only --steps 2 --warmup 1 > gpurun_out/h_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:extract_tiled -c 1 -o gpurun_out/prof_r01h_tiled -f python bench.py --kernel-only --steps 2 --warmup 1 > gpurun_out/h_ncu_full.log 2>&1
tail -2 gpurun_out/h_ncu_full.log
