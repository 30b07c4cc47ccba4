#!/bin/bash
# Gauss-point assembly with J^-1 / detJ evaluated once per element: parity subset + timings (affine pass off = every node general)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hex8 or ke or k5 or k6 or inverted or checked or config4_assembly or config3_full_vs or bad_connectivity or zero_rhs" > gpurun_out/ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ah_pytest.log | cut -c1-300
for f in 0 1; do
  FEA_ASSEMBLE_AFFINE=$f timeout 300 python tools/profile_kernels.py --only asm --hex 400 80 2>&1 | tail -1 | cut -c1-330
done
