#!/bin/bash
# L2 hint of the matrix stream at config 3 (85 MB matrix): evict-first / normal / evict-last
mkdir -p gpurun_out
for h in 0 1 2; do
  echo "FEA_TMA_L2=$h"
  for rep in 1 2; do
  FEA_TMA_L2=$h timeout 300 python tools/bench_configs.py 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['pcg_iterations'], d['ms'])"
  done
done
