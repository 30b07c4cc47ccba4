#!/bin/bash
# N GPUs: parity at N ranks, bench, config 5 distributed
mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/check_dist.py 96 16 --out gpurun_out/p_dist_check_n$N.json > gpurun_out/p_dist_check_n$N.log 2>&1
grep -E "DIST CHECK|\"ok\": false" gpurun_out/p_dist_check_n$N.log | cut -c1-500 | head -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/p_bench_n$N.json 2> gpurun_out/p_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/p_bench_n$N.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['solve']['pcg_iterations'], d['stages_ms_rank0'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['d2h_bytes_per_step'], d['clocks'])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 tools/bench_configs.py 5 --dist > gpurun_out/p_config5_n$N.json 2> gpurun_out/p_config5_n$N.err; echo "config5 rc=$?"
cat gpurun_out/p_config5_n$N.json | cut -c1-600
