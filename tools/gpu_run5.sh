#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_gnn.py -q -x ) > gpurun_out/r02_pytest5.log 2>&1
echo "exit $?" >> gpurun_out/r02_pytest5.log
tail -6 gpurun_out/r02_pytest5.log
timeout 300 python tools/persist_diag.py > gpurun_out/r02_persist_diag2.log 2>&1
E=15 timeout 300 python tools/persist_diag.py > gpurun_out/r02_persist_diag2_E15.log 2>&1
tail -4 gpurun_out/r02_persist_diag2.log; tail -4 gpurun_out/r02_persist_diag2_E15.log
GCBF_PERSISTENT=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r02_bench5_5launch.json 2> gpurun_out/r02_bench5_5launch.err
GCBF_PERSISTENT=0 GCBF_CHAIN8=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r02_bench5_5launch_old.json 2> gpurun_out/r02_bench5_old.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train --envs-per-gpu 15 > gpurun_out/r02_bench5_persist_E15.json 2> gpurun_out/r02_bench5_E15.err
python - <<'PY'
import json
for f in ("r02_bench5_5launch","r02_bench5_5launch_old","r02_bench5_persist_E15"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["config"]["us_per_env_step"], d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
