#!/bin/bash
B='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"])'
run() { python bench.py --kernel-only --workload $1 --steps 20 --warmup 5 2>/dev/null | python -c "$B"; }
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo -n "cfg4 fused: "; run cfg4
echo -n "cfg4 split: "; X360_SPLIT_SCORE=1 run cfg4
echo -n "cfg4 split again: "; X360_SPLIT_SCORE=1 run cfg4
echo -n "cfg4 noscore: "; run cfg4_noscore
