set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/h_cfg2.json 2> gpurun_out/h_cfg2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h_ref.json 2> gpurun_out/h_ref.err
for w in cfg1 cfg3 cfg4 cfg5; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_$w.json 2> gpurun_out/h_$w.err; done
python bench.py --kernel-only --steps 2 --warmup 1 > gpurun_out/h_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/h_launches.csv python bench.py --kernel-only --steps 2 --warmup 1 > gpurun_out/h_ncu_launch.log 2>&1
tail -c 300 gpurun_out/h_cfg2.json
