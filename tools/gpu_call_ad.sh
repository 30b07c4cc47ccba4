#!/bin/bash
# config 5 on N slabs with the peer-memory halo exchange: parity (check_dist), timing p2p vs nccl halo
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/check_dist.py 96 16 --out gpurun_out/ad_dist_check_n$N.json > gpurun_out/ad_dist_check_n$N.log 2>&1
grep -E "DIST CHECK|\"ok\": false|Error|error" gpurun_out/ad_dist_check_n$N.log | cut -c1-400 | head -8
for c in p2p nccl; do
FEA_DIST_COMM=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 tools/bench_configs.py 5 --dist > gpurun_out/ad_config5_n${N}_$c.json 2> gpurun_out/ad_config5_n${N}_$c.err; echo "config5 $c rc=$?"
grep config gpurun_out/ad_config5_n${N}_$c.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','seconds_slice_assemble_solve','seconds_solver_only','ms_per_iteration_solver_only','pcg_iterations','rel_residual_worst','status')})"
tail -2 gpurun_out/ad_config5_n${N}_$c.err | cut -c1-300
done
