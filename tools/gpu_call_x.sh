#!/bin/bash
# small / mid meshes: SpMV pipeline shape x L2 hint
mkdir -p gpurun_out
for size in "60 12" "100 20" "150 30" "200 40"; do
 for cfg in 2,2,3 3,3,2; do
  for h in 0 2; do
   FEA_TMA_L2=$h FEA_TMA_CFG=$cfg timeout 300 python tools/bench_configs.py 3 --c3 $size 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$size', '$cfg', 'L2=$h', d['dof'], d['pcg_iterations'], round(d['ms']['solve'],2), round(d['ms']['pcg_iteration']*1e3,2))"
  done
 done
done
