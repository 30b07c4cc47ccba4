#!/bin/bash
# round 2, call A (1 GPU): GPU test suite, N=1 bench, ncu slab traffic, sanitizers
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi -L > gpurun_out/a_gpus.txt; nproc >> gpurun_out/a_gpus.txt; free -g >> gpurun_out/a_gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/a_bench_n1.json
for A in 200 100 50; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
     -k regex:pcg_spmv_tma -c 3 --csv --log-file gpurun_out/a_ncu_slab_$A.csv python tools/profile_spmv.py $A 80 8 > gpurun_out/a_ncu_slab_$A.log 2>&1
done
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/a_sanitize_$tool.log python tools/sanitize.py > gpurun_out/a_sanitize_$tool.out 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/a_sanitize_$tool.log
done
