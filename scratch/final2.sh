#!/bin/bash
python bench.py > gpurun_out/j_bench_cfg2.json 2> gpurun_out/j_bench_cfg2.err; tail -1 gpurun_out/j_bench_cfg2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j_bench_ref.json 2> gpurun_out/j_bench_ref.err; tail -1 gpurun_out/j_bench_ref.json
for w in cfg1 cfg3 cfg4 cfg5; do python bench.py --workload $w --kernel-only > gpurun_out/j_bench_$w.json 2> gpurun_out/j_bench_$w.err; tail -1 gpurun_out/j_bench_$w.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j_ncu_launch.log 2>&1; tail -2 gpurun_out/j_ncu_launch.log
