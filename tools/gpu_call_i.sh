#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "k7 or chain or p2p_solver or pcg_single or beam" > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
tail -12 gpurun_out/i_pytest.log | cut -c1-400
for A in 50; do timeout 300 python tools/gated_probe.py $A 80 640 >> gpurun_out/i_gated_probe.log 2>&1; done
cat gpurun_out/i_gated_probe.log
