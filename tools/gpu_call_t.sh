#!/bin/bash
# config 5 on N slabs: CUDA-graph chunks vs eager loop
mkdir -p gpurun_out
N=${1:-2}
for g in 1 0; do
FEA_MULTI_GRAPH=$g timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$g tools/bench_configs.py 5 --dist > gpurun_out/t_config5_n${N}_g$g.json 2> gpurun_out/t_config5_n${N}_g$g.err; echo "config5 graph=$g rc=$?"
grep config gpurun_out/t_config5_n${N}_g$g.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','seconds_slice_assemble_solve','seconds_solver_only','ms_per_iteration_solver_only','pcg_iterations','graph')})"
done
