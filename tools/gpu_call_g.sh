#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "k7 or chain or p2p_solver or pcg_single" > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
tail -12 gpurun_out/g_pytest.log | cut -c1-300
for A in 50; do timeout 300 python tools/gated_probe.py $A 80 640 >> gpurun_out/g_gated_probe.log 2>&1; done
cat gpurun_out/g_gated_probe.log
for v in 1,0 1,1; do
  PROBE_ONLY=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_ncu_$v.csv python tools/gated_probe.py 50 80 48 > gpurun_out/g_probe_$v.log 2>&1
done
python - <<'PY'
import csv, collections
for v in ("1,0","1,1"):
    rows=[r for r in csv.reader(l for l in open(f"gpurun_out/g_ncu_{v}.csv") if not l.startswith("=="))]
    hdr=rows[0]; k=hdr.index("Kernel Name"); val=hdr.index("Metric Value")
    d=collections.defaultdict(list)
    for r in rows[1:]:
        if "pcg_" in r[k]: d[r[k].split("(")[0][-40:]].append(float(r[val].replace(",","")))
    for name,vals in d.items():
        vals=vals[len(vals)//4:]
        print(f"variant {v}  {name:40s} n={len(vals):4d} median {sorted(vals)[len(vals)//2]/1e3:8.2f} us  min {min(vals)/1e3:8.2f}")
PY
