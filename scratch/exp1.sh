#!/bin/bash
# experiment batch: baseline vs alternate builds of libx360 (cfg2 kernel-only + a parity subset)
B='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"])'
run() { python bench.py --kernel-only --workload $1 --steps 20 --warmup 5 2>/dev/null | python -c "$B"; }
for lib in "" $@; do
  if [ -n "$lib" ]; then export X360_LIB=$PWD/scratch/libs/$lib; else unset X360_LIB; fi
  echo "== lib=${lib:-default}"
  for w in cfg2 cfg1 cfg4; do echo -n "$w: "; run $w; done
  echo -n "cfg2 again: "; run cfg2
  if [ -n "$lib" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep or golden or hashes or full_size" 2>&1 | tail -3; fi
done
