#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/ai_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ai_pytest.log
tail -3 gpurun_out/ai_pytest.log | cut -c1-300
