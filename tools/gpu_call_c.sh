#!/bin/bash
# round 2, call C (2 GPUs): full GPU suite (incl. multi-rank parity), 2-rank check at a larger mesh, N=2 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -25 gpurun_out/c_pytest.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/check_dist.py 64 16 --out gpurun_out/c_dist_check_n2.json > gpurun_out/c_dist_check_n2.log 2>&1
grep -v '"ok": true' gpurun_out/c_dist_check_n2.log | tail -8 | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/c_bench_n2.json 2> gpurun_out/c_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['solve'], d['stages_ms_rank0'], d['roofline']['frac'], d['e2e']['ms_per_step'])
PY
