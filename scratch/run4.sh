timeout 900 python -m pytest tests -x -q -m gpu -k "unsharp or sharpens" 2>&1 | tail -3
python scripts/bench_unsharp.py 2>&1 | tail -1 | cut -c1-120
