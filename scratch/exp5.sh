#!/bin/bash
B='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"])'
run() { python bench.py --kernel-only --workload cfg2 --steps 20 --warmup 5 $@ 2>/dev/null | python -c "$B"; }
echo -n "default: "; run
for lv in 0 1 2 3; do echo -n "unit level $lv: "; X360_UNIT_LEVEL=$lv run; done
for fg in 8 16; do echo -n "FG=$fg: "; X360_FG=$fg run; done
echo -n "level1 FG=16: "; X360_UNIT_LEVEL=1 X360_FG=16 run
echo -n "level0 FG=16: "; X360_UNIT_LEVEL=0 X360_FG=16 run
echo -n "level0 FG=8: "; X360_UNIT_LEVEL=0 X360_FG=8 run
echo -n "frames=64: "; run --frames 64
