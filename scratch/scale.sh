#!/bin/bash
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/k_bench_cfg2_${N}gpu.json 2> gpurun_out/k_bench_cfg2_${N}gpu.err
tail -1 gpurun_out/k_bench_cfg2_${N}gpu.json
