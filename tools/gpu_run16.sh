#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
timeout 900 python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/r02_bench_config5_final.json 2> gpurun_out/r02_bench_config5_final.err
timeout 900 python bench.py --config 5 --steps 3 --warmup 3 --envs-per-gpu 64 --no-cpu-baseline --no-train > gpurun_out/r02_bench_config5_E64_final.json 2> gpurun_out/r02_bench_config5_E64_final.err
python - <<'PY'
import json
for f in ("r02_bench_final","r02_bench_config5_final","r02_bench_config5_E64_final"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); r=d["roofline"]
        print(f, d["value"], d["config"]["us_per_env_step"], r["bound"], round(r["frac"],4), r.get("traffic"), r.get("traffic_source","")[:60], [round(k["us"],1) for k in (r.get("step_kernels") or r["five_launch_path"]["step_kernels"])])
    except Exception as e: print(f, "ERR", e)
PY
