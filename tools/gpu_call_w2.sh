#!/bin/bash
# fused PCG at config 3: pipeline shapes
mkdir -p gpurun_out
for cfg in 3,3,2 2,2,3 2,4,2 4,4,1 3,3,1 2,2,2 1,2,3 4,8,1 3,6,1; do
  FEA_TMA_CFG=$cfg timeout 300 python tools/bench_configs.py 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', d['pcg_iterations'], round(d['ms']['solve'],2), round(d['ms']['pcg_iteration']*1e3,2))"
done
