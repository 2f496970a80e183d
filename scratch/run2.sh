timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
B='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"].get("tile_modes"))'
echo "cfg2 default"; python bench.py --kernel-only --steps 20 --warmup 3 2>/dev/null | python -c "$B"
echo "cfg2 m5"; X360_LIB=$PWD/360extractor_b200/lib/libx360_m5.so python bench.py --kernel-only --steps 20 --warmup 3 2>/dev/null | python -c "$B"
echo cfg4; python bench.py --kernel-only --workload cfg4 --steps 10 --warmup 3 2>/dev/null | python -c "$B"
echo cfg1; python bench.py --kernel-only --workload cfg1 --steps 10 --warmup 3 2>/dev/null | python -c "$B"
