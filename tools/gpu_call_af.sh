#!/bin/bash
# end of round 2: full GPU suite + smoke on the final tree, then one ncu --set full capture of the persistent PCG kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/af_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/af_pytest.log
tail -4 gpurun_out/af_pytest.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pcg_fused -c 1 -o gpurun_out/af_fused -f python tools/bench_configs.py 3 > gpurun_out/af_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/af_fused.ncu-rep --page raw --csv > gpurun_out/af_fused_raw.csv 2>/dev/null; ls -la gpurun_out/af_fused* | cut -c1-120
