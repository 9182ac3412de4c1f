#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py tests/test_gpu_qp.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py tests/test_gpu_gnn.py -x -q -m gpu > gpurun_out/r02_pytest_train17.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_train17.log
timeout 600 python bench.py --train-only --T 64 > gpurun_out/r02_train_only17.json 2> gpurun_out/r02_train_only17.err
cat gpurun_out/r02_train_only17.json | cut -c1-400
GCBF_TRAIN_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches17.csv python bench.py --train-only --T 8 > gpurun_out/r02_train_ncu17.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open("gpurun_out/r02_train_launches17.csv") if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.OrderedDict()
data=rows[1:]
# keep the last 190 launches (one eager step at the end)
for r in data[-175:]:
    v=float(r[vi].replace(",","")); u=r[ui]
    v = v/1e3 if u in ("ns","nsecond") else (v*1e3 if u in ("ms","msecond") else v)
    k=r[ki][:60]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{k:60s} {a[0]:4d} {a[1]:9.1f} us {100*a[1]/tot:5.1f}%")
print("total", tot)
PY
