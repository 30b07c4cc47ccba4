#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent or pcg or k5 or config3_full_vs or checked" 2>&1 | tail -2
for rep in 1 2; do
timeout 300 python tools/bench_configs.py 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('config3', d['pcg_iterations'], round(d['ms']['solve'],2), round(d['ms']['pcg_iteration']*1e3,2))"
done
timeout 300 python tools/bench_configs.py 3 --c3 60 12 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('60x12', d['pcg_iterations'], round(d['ms']['solve'],2), round(d['ms']['pcg_iteration']*1e3,2))"
