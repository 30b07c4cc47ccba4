#!/bin/bash
# N=1 bench (no cpu baseline) to see the iteration count with the affine-assembled K + ncu of the assembly kernels
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s_bench_n1.json 2> gpurun_out/s_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/s_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['solve'], d.get('stages_ms'), d['roofline']['frac'], d['e2e'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"assemble_hex8|affine_geometry" -c 3 -o gpurun_out/s_asm -f python tools/profile_kernels.py --once --only asm --hex 400 80 > gpurun_out/s_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/s_asm.ncu-rep --page raw --csv > gpurun_out/s_asm_raw.csv 2>/dev/null
