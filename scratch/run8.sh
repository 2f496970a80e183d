python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/b8.json 2> gpurun_out/b8.err
tail -1 gpurun_out/b8.json | cut -c1-400
