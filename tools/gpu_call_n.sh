#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/check_dist.py 64 16 --out gpurun_out/n_dist_check_n$N.json > gpurun_out/n_dist_check_n$N.log 2>&1
grep -E "DIST CHECK|\"ok\": false|Error|error" gpurun_out/n_dist_check_n$N.log | cut -c1-600 | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 tools/bench_configs.py 5 --dist > gpurun_out/n_config5_n$N.json 2> gpurun_out/n_config5_n$N.err; echo "config5 rc=$?"
cat gpurun_out/n_config5_n$N.json; tail -3 gpurun_out/n_config5_n$N.err
