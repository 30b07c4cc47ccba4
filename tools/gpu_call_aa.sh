#!/bin/bash
# round 2, session 2, 1 GPU: full GPU suite, smoke, N=1 bench, configs 2/3/5, assembly timings
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/aa_pytest.log
tail -6 gpurun_out/aa_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/aa_bench_n1.json 2> gpurun_out/aa_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/aa_bench_n1.json').read().strip().splitlines()[-1])
print("ours:", d['value'], d['ms_per_step'], d['solve'], d['stages_ms'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])
PY
timeout 900 python tools/bench_configs.py 2 3 5 > gpurun_out/aa_configs.jsonl 2> gpurun_out/aa_configs.err; cat gpurun_out/aa_configs.jsonl | cut -c1-700
timeout 300 python tools/profile_kernels.py --only asm --hex 400 80 2>&1 | tail -1 | cut -c1-400
