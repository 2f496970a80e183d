#!/bin/bash
B='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"])'
run() { python bench.py --kernel-only --workload $1 --steps 20 --warmup 5 2>/dev/null | python -c "$B"; }
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
echo -n "cfg2 nb3 (default): "; run cfg2
echo -n "cfg2 nb3 again: "; run cfg2
echo -n "cfg2 nb2: "; X360_LIB=$PWD/scratch/libs/libx360_nb2.so run cfg2
echo -n "cfg2 nb3 rows=48 (6 CTAs): "; X360_TILED_ROWS=48 run cfg2
echo -n "cfg2 nb3 bw=128: "; X360_TILED_BW=128 run cfg2
echo -n "cfg2 nb3 bw=256: "; X360_TILED_BW=256 run cfg2
for w in cfg1 cfg3 cfg4 cfg5; do echo -n "$w: "; run $w; done
