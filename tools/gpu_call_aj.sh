#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent or pcg_vs or k5" 2>&1 | tail -2
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
