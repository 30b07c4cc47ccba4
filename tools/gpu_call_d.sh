#!/bin/bash
# round 2, call D (1 GPU): new tests, gated-vs-plain SpMV A/B on slabs, configs bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "k8 or relax or mesh_builders or euler or config2" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -30 gpurun_out/d_pytest.log | cut -c1-300
for A in 200 50; do timeout 300 python tools/gated_probe.py $A 80 640 >> gpurun_out/d_gated_probe.log 2>&1; done
cat gpurun_out/d_gated_probe.log
