python bench.py --kernel-only --workload cfg3 --steps 3 --warmup 1 > gpurun_out/lz_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:extract_tiled -c 1 -o gpurun_out/prof_lz_cfg3 -f python bench.py --kernel-only --workload cfg3 --steps 2 --warmup 1 > gpurun_out/lz_ncu.log 2>&1
tail -1 gpurun_out/lz_plain.log | cut -c1-200
