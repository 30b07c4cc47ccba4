#!/bin/bash
# round 2, call F (1 GPU): per-kernel durations of the gated vs plain PCG kernels under ncu (relative comparison)
mkdir -p gpurun_out
for v in 1,0 1,1; do
  PROBE_ONLY=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_ncu_$v.csv python tools/gated_probe.py 50 80 48 > gpurun_out/f_probe_$v.log 2>&1
done
python - <<'PY'
import csv, collections
for v in ("1,0","1,1"):
    rows=[r for r in csv.reader(l for l in open(f"gpurun_out/f_ncu_{v}.csv") if not l.startswith("=="))]
    hdr=rows[0]; k=hdr.index("Kernel Name"); val=hdr.index("Metric Value")
    d=collections.defaultdict(list)
    for r in rows[1:]:
        d[r[k].split("(")[0][-60:]].append(float(r[val].replace(",","")))
    print("variant",v)
    for name,vals in d.items():
        vals=vals[len(vals)//4:]
        print(f"  {name:60s} n={len(vals):4d} median {sorted(vals)[len(vals)//2]/1e3:8.2f} us  min {min(vals)/1e3:8.2f}")
PY
