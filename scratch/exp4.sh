#!/bin/bash
B='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"])'
run() { python bench.py --kernel-only --workload $1 --steps 20 --warmup 5 2>/dev/null | python -c "$B"; }
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
echo -n "cfg2 default (trimmed): "; run cfg2
echo -n "cfg2 default again: "; run cfg2
export X360_LIB=$PWD/scratch/libs/libx360_pred3.so
echo -n "cfg2 pred3: "; run cfg2
echo -n "cfg2 pred3 again: "; run cfg2
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep or golden or hashes or full_size" 2>&1 | tail -3
