#!/bin/bash
# final pass of round 2 (after the folded train step): parity suite, smoke, traffic capture of this build, bench lines
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_multi.py ) > gpurun_out/r02_pytest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_pytest_final.log
tail -5 gpurun_out/r02_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; tail -2 gpurun_out/r02_smoke_final.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_persist_kernel -s 4 -c 1 -o gpurun_out/r02_persist_full -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r02_ncu_full_persist.log 2>&1
ncu -i gpurun_out/r02_persist_full.ncu-rep --page raw --csv > gpurun_out/r02_persist_full_raw.csv 2>/dev/null
python tools/ncu_traffic.py gpurun_out/r02_persist_full_raw.csv > gpurun_out/r02_traffic_tool.log 2>&1; cp profiles/r02_traffic.json gpurun_out/r02_traffic.json
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_config4_final.json 2> gpurun_out/r02_bench_config4_final.err
timeout 600 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_config5_final.json 2> gpurun_out/r02_bench_config5_final.err
GCBF_TRAIN_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches_final.csv python bench.py --train-only --T 8 > gpurun_out/r02_train_ncu_final.log 2>&1
GCBF_TRAIN_GRAPH=0 timeout 600 ncu --set full --clock-control none -k regex:"gemm_tc_kernel|gemm_tn_tc_kernel" -s 40 -c 16 -o gpurun_out/r02_train_gemm_full -f python bench.py --train-only --T 8 > gpurun_out/r02_train_ncu_full.log 2>&1
ncu -i gpurun_out/r02_train_gemm_full.ncu-rep --page raw --csv > gpurun_out/r02_train_gemm_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02_persist_full.ncu-rep gpurun_out/r02_train_gemm_full.ncu-rep
python - <<'PY'
import json
for f in ("r02_bench_final","r02_bench_config4_final","r02_bench_config5_final"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); r=d["roofline"]
        print(f, d["value"], d["e2e"]["value"], d.get("gpu_launches"), (d.get("train_step") or {}).get("ms_per_minibatch"), r.get("traffic"), round(r["frac"],4))
    except Exception as e: print(f, "ERR", e)
PY
