#!/bin/bash
mkdir -p gpurun_out
PROBE_ONLY=0,1 timeout 300 python tools/gated_probe.py 200 80 320 > gpurun_out/l_probe.log 2>&1
FEA_P2P_FAKE_TILES=1 PROBE_ONLY=0,1 timeout 300 python tools/gated_probe.py 200 80 320 >> gpurun_out/l_probe.log 2>&1
PROBE_ONLY=0,0 timeout 300 python tools/gated_probe.py 200 80 320 >> gpurun_out/l_probe.log 2>&1
cat gpurun_out/l_probe.log
