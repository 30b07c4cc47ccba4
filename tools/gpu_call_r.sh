#!/bin/bash
# ncu --set full of the two hex8 assembly kernels at 400x80x80 (one launch each)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex8 -c 2 -o gpurun_out/r_asm -f python tools/profile_kernels.py --once --only asm --hex 400 80 > gpurun_out/r_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r_ncu.log | cut -c1-300
ncu -i gpurun_out/r_asm.ncu-rep --page raw --csv > gpurun_out/r_asm_raw.csv 2>/dev/null
ncu -i gpurun_out/r_asm.ncu-rep --page source --csv -k regex:affine > gpurun_out/r_asm_src.csv 2>/dev/null
ls -la gpurun_out/r_asm*
