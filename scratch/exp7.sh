#!/bin/bash
B='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["tile_modes"], d["plan"])'
run() { python bench.py --kernel-only --workload $1 --steps 20 --warmup 5 2>/dev/null | python -c "$B"; }
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo -n "cfg4 raw: "; run cfg4
echo -n "cfg4 planes: "; X360_NO_RAW=1 run cfg4
echo -n "cfg4 raw th=32: "; X360_TILED_TH=32 run cfg4
echo -n "cfg1 raw: "; run cfg1
echo -n "cfg2 raw: "; run cfg2
python scripts/full_parity.py gpurun_out/r01k_full_parity.json 2>&1 | tail -2
