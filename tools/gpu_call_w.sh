#!/bin/bash
# config 3: sweep of the SpMV pipeline shape with the matrix kept in L2 (FEA_TMA_L2=2), and the in-graph SpMV floor
mkdir -p gpurun_out
export FEA_TMA_L2=2
for cfg in 2,2,3 1,2,3 1,1,6 2,2,2 3,3,2 1,3,2 2,4,1 4,4,1 1,2,4 1,1,5; do
  FEA_TMA_CFG=$cfg timeout 300 python tools/bench_configs.py 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', d['pcg_iterations'], round(d['ms']['solve'],2), round(d['ms']['spmv']*1e3,1), round(d['ms']['pcg_iteration']*1e3,2))"
done
for h in 0 2; do FEA_TMA_L2=$h timeout 300 python tools/experiments/spmv_in_graph.py 100 20 2>&1 | tail -1; done
FEA_TMA_L2=0 timeout 300 python tools/experiments/spmv_in_graph.py 400 80 100 2>&1 | tail -1
