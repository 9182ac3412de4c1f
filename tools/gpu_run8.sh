#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_rollout.py -q -x -k "persistent" ) > gpurun_out/r02_pytest8.log 2>&1
echo "exit $?" >> gpurun_out/r02_pytest8.log; tail -4 gpurun_out/r02_pytest8.log
GCBF_PERSIST_COOP=1 timeout 300 python tools/persist_diag.py > gpurun_out/r02_persist_diag_coop.log 2>&1; tail -3 gpurun_out/r02_persist_diag_coop.log
# sanitizers on the persistent kernel (cluster mode and pair mode), small scene
for mode in 0 1; do
  for tool in memcheck racecheck; do
    GCBF_PERSIST_SOFT=$mode GCBF_SANITIZE_PERSISTENT=1 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py DoubleIntegrator > gpurun_out/r02_sanitizer_${tool}_persist_mode${mode}.log 2>&1
    echo "exit $?" >> gpurun_out/r02_sanitizer_${tool}_persist_mode${mode}.log
    tail -3 gpurun_out/r02_sanitizer_${tool}_persist_mode${mode}.log
  done
done
# configs 4 and 5 at their per-GPU shard sizes, and config 5 sweep
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r02_bench_config4.json 2> gpurun_out/r02_bench_config4.err
timeout 900 python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/r02_bench_config5.json 2> gpurun_out/r02_bench_config5.err
for e in 16 32 64; do
  timeout 900 python bench.py --config 5 --steps 3 --warmup 3 --envs-per-gpu $e --no-cpu-baseline --no-train > gpurun_out/r02_bench_config5_E$e.json 2> gpurun_out/r02_bench_config5_E$e.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench_config*.json")):
    try:
        d=json.load(open(f)); print(f, d["value"], d["config"]["us_per_env_step"], d["gpu_launches"], (d.get("train_step") or {}).get("ms_per_minibatch"))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-300:])
PY
