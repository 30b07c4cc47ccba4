#!/bin/bash
mkdir -p gpurun_out
for v in 1,0 1,1; do
  PROBE_ONLY=$v timeout 300 ncu --set full --import-source on --clock-control none -k regex:pcg_spmv_tma -s 20 -c 1 -f -o gpurun_out/h_spmv_$v python tools/gated_probe.py 50 80 48 > gpurun_out/h_probe_$v.log 2>&1
  PROBE_ONLY=$v timeout 300 ncu --set full --import-source on --clock-control none -k regex:pcg_cgcg -s 20 -c 1 -f -o gpurun_out/h_cgcg_$v python tools/gated_probe.py 50 80 48 >> gpurun_out/h_probe_$v.log 2>&1
done
ls -la gpurun_out/h_*
