#!/bin/bash
# round 2, call E (1 GPU): EB chain solver tests, gated-vs-plain SpMV A/B after the PeerKey refactor
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "k7 or chain or config2 or p2p_solver or pcg" > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -30 gpurun_out/e_pytest.log | cut -c1-300
for A in 200 50; do timeout 300 python tools/gated_probe.py $A 80 640 >> gpurun_out/e_gated_probe.log 2>&1; done
cat gpurun_out/e_gated_probe.log
