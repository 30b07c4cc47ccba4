#!/bin/bash
# round 2, call B (2 GPUs): multi-rank parity + config-4 full vs oracle fixture + N=2 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_rank or config4_full" > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/check_dist.py 64 16 --out gpurun_out/b_dist_check_n2.json > gpurun_out/b_dist_check_n2.log 2>&1
tail -3 gpurun_out/b_dist_check_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/b_bench_n2.json 2> gpurun_out/b_bench_n2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/b_bench_n2.json; tail -5 gpurun_out/b_bench_n2.err
