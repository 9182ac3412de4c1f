#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py tests/test_gpu_gnn.py tests/test_gpu_fullsize.py tests/test_gpu_qp.py -x -q -m gpu > gpurun_out/r02_pytest_29.log 2>&1
rc1=$?; echo "pytest rc=$rc1"; tail -3 gpurun_out/r02_pytest_29.log
timeout 200 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k "persistent" > gpurun_out/r02_pytest_29b.log 2>&1
rc2=$?; echo "pytest rollout rc=$rc2"; tail -2 gpurun_out/r02_pytest_29b.log
timeout 300 python bench.py --train-only 2> gpurun_out/r02_train_only29.err | cut -c1-200
if [ $rc1 -ne 0 ] || [ $rc2 -ne 0 ]; then echo "tests failed: stopping"; exit 1; fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_persist_kernel -s 4 -c 1 -o gpurun_out/r02_persist_full -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r02_ncu_full_persist.log 2>&1
ncu -i gpurun_out/r02_persist_full.ncu-rep --page raw --csv > gpurun_out/r02_persist_full_raw.csv 2>/dev/null
python tools/ncu_traffic.py gpurun_out/r02_persist_full_raw.csv > gpurun_out/r02_traffic_tool.log 2>&1; cp profiles/r02_traffic.json gpurun_out/r02_traffic.json
rm -f gpurun_out/r02_persist_full.ncu-rep
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_final.json") if l.startswith("{")][-1]); r=d["roofline"]
print(d["value"], d["e2e"]["value"], d["train_step"]["ms_per_minibatch"], r.get("traffic"), round(r["frac"],4))
PY
