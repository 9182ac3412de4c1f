#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_21.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_21.log
GCBF_TRAIN_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches21.csv python bench.py --train-only --T 8 > gpurun_out/r02_train_ncu21.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_21.json 2> gpurun_out/r02_bench_21.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_21.json") if l.startswith("{")][-1])
print(d["value"], d["e2e"]["value"], d["train_step"])
PY
