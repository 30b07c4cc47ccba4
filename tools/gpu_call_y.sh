#!/bin/bash
# persistent fused PCG: parity subset (bounded), then config 3 and other sizes fused on/off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pcg or k5 or k6 or config3 or checked or hex8_pattern or cubebeam" > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/y_pytest.log | cut -c1-300
for f in 1 0; do
 for size in "60 12" "100 20" "150 30"; do
  FEA_PCG_FUSED=$f timeout 300 python tools/bench_configs.py 3 --c3 $size 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fused=$f', '$size', d['dof'], d['pcg_iterations'], d['rel_residual'], round(d['ms']['solve'],2), round(d['ms']['pcg_iteration']*1e3,2))"
 done
done
