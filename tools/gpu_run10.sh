#!/bin/bash
mkdir -p gpurun_out
( time timeout 2700 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_multi.py ) > gpurun_out/r02_pytest10.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_pytest10.log
tail -8 gpurun_out/r02_pytest10.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err
timeout 600 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r02_bench10_config5.json 2> gpurun_out/r02_bench10_config5.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke10.log 2>&1; tail -2 gpurun_out/r02_smoke10.log
python - <<'PY'
import json
for f in ("r02_bench10","r02_bench10_config5"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["e2e"]["value"], d["config"]["us_per_env_step"], d["gpu_launches"], (d.get("train_step") or {}).get("ms_per_minibatch"), d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
