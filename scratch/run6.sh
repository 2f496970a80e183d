python scripts/bench_unsharp.py --n 64 --steps 2 --warmup 1 > gpurun_out/us_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:unsharp_tiled -c 1 -o gpurun_out/prof_unsharp -f python scripts/bench_unsharp.py --n 64 --steps 2 --warmup 1 > gpurun_out/us_ncu.log 2>&1
tail -2 gpurun_out/us_ncu.log
