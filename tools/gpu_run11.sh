#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_rollout.py tests/test_gpu_train.py tests/test_reference_goldens.py tests/test_gpu_trainer.py -q --maxfail=10 ) > gpurun_out/r02_pytest11.log 2>&1
echo "exit $?" >> gpurun_out/r02_pytest11.log; tail -6 gpurun_out/r02_pytest11.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke11.log 2>&1; tail -2 gpurun_out/r02_smoke11.log
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench11_config4.json 2> gpurun_out/r02_bench11_config4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench11_config4.json")); print(d["value"], d["config"]["us_per_env_step"], d["gpu_launches"], d["train_step"]["ms_per_minibatch"])
PY
