#!/bin/bash
# 8 GPUs: parity at 8 ranks, N=8 bench, iteration timeline
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/check_dist.py 96 16 --out gpurun_out/m_dist_check_n$N.json > gpurun_out/m_dist_check_n$N.log 2>&1
grep -E "DIST CHECK|\"ok\": false" gpurun_out/m_dist_check_n$N.log | cut -c1-500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/m_bench_n$N.json 2> gpurun_out/m_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/m_bench_n$N.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['solve']['pcg_iterations'], d['stages_ms_rank0'], d['roofline']['frac'], d['e2e'])
PY
FEA_P2P_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/p2p_debug.py 400 80 400 2>&1 | grep -E "rep|rank [03] it 20[3-4]|plain" | cut -c1-200
