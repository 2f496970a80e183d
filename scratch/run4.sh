timeout 900 python -m pytest tests -x -q -m gpu -k "unsharp or sharpens" 2>&1 | tail -15
