#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent or p2p_solver or pcg" > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/ab_pytest.log | cut -c1-300
