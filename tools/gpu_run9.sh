#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_graph.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py tests/test_gpu_reset.py -q --maxfail=5 ) > gpurun_out/r02_pytest9.log 2>&1
echo "exit $?" >> gpurun_out/r02_pytest9.log; tail -5 gpurun_out/r02_pytest9.log
( timeout 600 python -m pytest tests/test_gpu_rollout.py -q -k "LinearDrone" ) > gpurun_out/r02_pytest9b.log 2>&1; tail -2 gpurun_out/r02_pytest9b.log
timeout 900 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r02_bench9_config5.json 2> gpurun_out/r02_bench9_config5.err
timeout 900 python bench.py --config 5 --steps 3 --warmup 3 --envs-per-gpu 64 --no-cpu-baseline --no-train > gpurun_out/r02_bench9_config5_E64.json 2> gpurun_out/r02_bench9_config5_E64.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench9*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f, d["value"], d["config"]["us_per_env_step"], [round(k["us"],1) for k in r["step_kernels"]], r["bound"], r["frac"])
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-300:])
PY
