#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/check_dist.py 64 16 --out gpurun_out/j_dist_check_n2.json > gpurun_out/j_dist_check_n2.log 2>&1
grep -E "DIST CHECK|\"ok\": false" gpurun_out/j_dist_check_n2.log | cut -c1-700
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/j_bench_n2.json 2> gpurun_out/j_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/j_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['solve']['pcg_iterations'], d['stages_ms_rank0'], d['roofline']['frac'], d['e2e']['ms_per_step'])
PY
FEA_P2P_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/p2p_debug.py 400 80 400 2>&1 | grep -E "rep|rank 0 it 20[3-4]|plain" | cut -c1-200
