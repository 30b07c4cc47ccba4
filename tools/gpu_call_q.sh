#!/bin/bash
# round 2, session 2: affine assembly pass + bitonic symbolic pass: parity subset, timings on/off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hex8 or k5 or k6 or config4_assembly or inverted or checked or beam_and_truss or config3_full_vs" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/q_pytest.log
for f in 1 0; do
  FEA_ASSEMBLE_AFFINE=$f timeout 300 python tools/profile_kernels.py --only asm --hex 400 80 2>&1 | tail -1 | cut -c1-400
done
FEA_ASSEMBLE_AFFINE=1 timeout 300 python tools/profile_kernels.py --only asm --hex 100 20 2>&1 | tail -1 | cut -c1-400
