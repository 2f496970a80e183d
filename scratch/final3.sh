#!/bin/bash
python bench.py > gpurun_out/k_bench_cfg2.json 2> gpurun_out/k_bench_cfg2.err; tail -1 gpurun_out/k_bench_cfg2.json | cut -c1-160
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/k_bench_ref.json 2> gpurun_out/k_bench_ref.err; tail -1 gpurun_out/k_bench_ref.json | cut -c1-160
for w in cfg1 cfg3 cfg4 cfg5; do python bench.py --workload $w --kernel-only > gpurun_out/k_bench_$w.json 2> gpurun_out/k_bench_$w.err; tail -1 gpurun_out/k_bench_$w.json | cut -c1-160; done
python scripts/bench_unsharp.py 2>/dev/null | tail -1 > gpurun_out/k_bench_unsharp.json; cut -c1-200 gpurun_out/k_bench_unsharp.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/k_ncu_launch.log 2>&1; tail -1 gpurun_out/k_ncu_launch.log | cut -c1-100
ncu --set full --clock-control none --import-source on -k regex:extract_tiled -c 1 -o gpurun_out/prof_r01k_raw python bench.py --kernel-only --workload cfg2 --steps 2 --warmup 1 > gpurun_out/k_ncu.log 2>&1; tail -1 gpurun_out/k_ncu.log
