python scripts/bench_unsharp.py 2>&1 | tail -1
python scripts/bench_unsharp.py --d 1920 --n 96 2>&1 | tail -1
X360_UNSHARP_GENERIC=1 python scripts/bench_unsharp.py --n 32 --steps 5 2>&1 | tail -1
