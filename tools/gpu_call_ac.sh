#!/bin/bash
# same-box A/B of the multi-GPU kernels (forced on one GPU, the slab of a rank at N = 8): session-start library vs now;
# then config 3 with the current persistent kernel.  The old library is not kept in the tree; rebuild it with
#   git archive e050de3 fea_b200 include | tar -x -C /tmp/oldsrc && (cd /tmp/oldsrc && python -c "from fea_b200 import build; build.build_library(force=True)")
#   mkdir -p tools/experiments/_ab && cp /tmp/oldsrc/fea_b200/csrc/libfea_b200.so tools/experiments/_ab/libfea_b200_e050de3.so
mkdir -p gpurun_out
OLD=tools/experiments/_ab/libfea_b200_e050de3.so
for rep in 1 2; do
  for lib in new old; do
    if [ $lib = old ]; then export FEA_LIB_PATH=$PWD/$OLD; else unset FEA_LIB_PATH; fi
    echo "== $lib (rep $rep)"
    FEA_P2P_FAKE_TILES=1 PROBE_ONLY=1,1 timeout 300 python tools/gated_probe.py 50 80 3000 2>&1 | tail -1
    FEA_P2P_FAKE_TILES=1 PROBE_ONLY=1,0 timeout 300 python tools/gated_probe.py 50 80 3000 2>&1 | tail -1
  done
done
unset FEA_LIB_PATH
for rep in 1 2; do
timeout 300 python tools/bench_configs.py 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('config3', d['pcg_iterations'], round(d['ms']['solve'],2), round(d['ms']['pcg_iteration']*1e3,2))"
done
timeout 300 python tools/bench_configs.py 3 --c3 60 12 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('60x12', d['pcg_iterations'], round(d['ms']['solve'],2), round(d['ms']['pcg_iteration']*1e3,2))"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent or pcg" 2>&1 | tail -2
