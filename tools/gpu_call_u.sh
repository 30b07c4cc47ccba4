#!/bin/bash
# programmatic dependent launch in the single-GPU PCG: parity subset + config 3 timings on/off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pcg or k5 or config3_full_vs or p2p_solver_single" > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/u_pytest.log
for f in 1 0; do
  echo "PDL=$f"
  FEA_PCG_PDL=$f timeout 300 python tools/bench_configs.py 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['pcg_iterations'], d['ms'])"
  FEA_PCG_PDL=$f timeout 300 python tools/bench_configs.py 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['pcg_iterations'], d['ms'])"
done
