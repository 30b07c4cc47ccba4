#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/check_dist.py 64 16 --out gpurun_out/ae_dist_check_n$N.json > gpurun_out/ae_dist_check_n$N.log 2>&1
grep -E "DIST CHECK|\"ok\": false|Error|error" gpurun_out/ae_dist_check_n$N.log | cut -c1-600 | head -8
python - <<PY
import json,re
for line in open('gpurun_out/ae_dist_check_n$N.log'):
    if '"variant"' in line and '"rank": 0' in line:
        try:
            d=json.loads(line); print(d['variant'][:60], d.get('iterations'), d.get('iterations_1gpu'), d.get('u_err'), d.get('ok'))
        except Exception as e: pass
PY
