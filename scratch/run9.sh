timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
B='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"])'
echo "cfg2 col"; python bench.py --kernel-only --steps 20 --warmup 3 2>/dev/null | python -c "$B"
echo "cfg2 nocol"; X360_NO_COL_U0=1 python bench.py --kernel-only --steps 20 --warmup 3 2>/dev/null | python -c "$B"
echo "cfg4_ring"; python bench.py --kernel-only --workload cfg4_ring --steps 10 --warmup 3 2>/dev/null | python -c "$B"
