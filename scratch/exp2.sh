#!/bin/bash
B='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"])'
run() { python bench.py --kernel-only --workload $1 --steps 20 --warmup 5 2>/dev/null | python -c "$B"; }
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
echo -n "cfg2 raw: "; run cfg2
echo -n "cfg2 raw again: "; run cfg2
echo -n "cfg2 planes: "; X360_NO_RAW=1 run cfg2
for r in 40 48; do for bw in 128 192 256; do echo -n "cfg2 raw rows=$r bw=$bw: "; X360_TILED_ROWS=$r X360_TILED_BW=$bw run cfg2; done; done
