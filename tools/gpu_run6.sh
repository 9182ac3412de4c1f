#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_rollout.py -q -x -k "persistent" ) > gpurun_out/r02_pytest6.log 2>&1
echo "exit $?" >> gpurun_out/r02_pytest6.log
tail -5 gpurun_out/r02_pytest6.log
timeout 300 python tools/persist_diag.py > gpurun_out/r02_persist_diag3_soft.log 2>&1
tail -22 gpurun_out/r02_persist_diag3_soft.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench6.json")); print(d["value"], d["config"]["us_per_env_step"], d["gpu_launches"]); print(d["roofline"].get("phases_us"))
PY
tail -3 gpurun_out/r02_bench6.err
