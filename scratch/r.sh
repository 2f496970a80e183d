B='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"])'
for th in 16 32; do for w in cfg1 cfg4 cfg4_noscore; do echo "$w th=$th"; X360_TILED_TH=$th python bench.py --kernel-only --workload $w --steps 10 --warmup 3 2>/dev/null | python -c "$B"; done; done
