#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1" >> gpurun_out/k_debug.log; env $1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/p2p_debug.py 400 80 400 >> gpurun_out/k_debug.log 2>&1; }
run "FEA_X=1"
grep -E "==|rep|p2p dbg|plain|Error|error|span" gpurun_out/k_debug.log | sed 's/\[p2p dbg/\n[p2p dbg/g' | grep -E "==|rep|rank 0 it 20[3-4]|plain|rror|span" | cut -c1-200
