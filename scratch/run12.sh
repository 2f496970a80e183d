timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
B='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"])'
for w in cfg2 cfg3 cfg5 cfg1; do echo $w; python bench.py --kernel-only --workload $w --steps 20 --warmup 3 2>/dev/null | python -c "$B"; done
