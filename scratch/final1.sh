#!/bin/bash
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/full_parity.py gpurun_out/r01j_full_parity.json 2>&1 | tail -9
