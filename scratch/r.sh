python bench.py --kernel-only --steps 2 --warmup 1 > gpurun_out/i_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:extract_tiled -c 1 -o gpurun_out/prof_hipad -f python bench.py --kernel-only --steps 2 --warmup 1 > gpurun_out/i_ncu.log 2>&1
tail -1 gpurun_out/i_ncu.log
