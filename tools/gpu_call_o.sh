#!/bin/bash
# 1 GPU: full GPU suite, reference arm at the driver's step counts, N=1 bench, configs 2/3/5
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log
tail -15 gpurun_out/o_pytest.log | cut -c1-300
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/o_bench_reference.json 2> gpurun_out/o_bench_reference.err ) 2> gpurun_out/o_bench_reference.time; echo "reference rc=$?"; tail -3 gpurun_out/o_bench_reference.time
timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/o_bench_n1.json 2> gpurun_out/o_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
r=json.loads(open('gpurun_out/o_bench_reference.json').read().strip().splitlines()[-1])
d=json.loads(open('gpurun_out/o_bench_n1.json').read().strip().splitlines()[-1])
print("reference:", r['value'], r['ms_per_step'], r['cpu_baseline']['cores'], r['cpu_baseline']['sample'][:400])
print("ours:", d['value'], d['ms_per_step'], d['solve'], d['stages_ms'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])
print("config equal:", r['config']==d['config'])
PY
timeout 900 python tools/bench_configs.py 2 3 5 > gpurun_out/o_configs.jsonl 2> gpurun_out/o_configs.err; cat gpurun_out/o_configs.jsonl | cut -c1-900
