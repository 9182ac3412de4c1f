#!/bin/bash
# final pass of the round: full parity suite, smoke, bench lines, ncu launch lists + full capture of the dominant kernel
mkdir -p gpurun_out
( time timeout 2700 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_multi.py ) > gpurun_out/r02_pytest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_pytest_final.log
tail -6 gpurun_out/r02_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; tail -2 gpurun_out/r02_smoke_final.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r02_ncu_launches_final.log 2>&1
GCBF_PERSISTENT=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/r02_launches_5launch_final_T16.csv \
   python bench.py --steps 2 --warmup 3 --T 16 --no-cpu-baseline --no-train > gpurun_out/r02_ncu_launches_5launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_persist_kernel -s 4 -c 1 -o gpurun_out/r02_persist_full -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r02_ncu_full_persist.log 2>&1
ncu -i gpurun_out/r02_persist_full.ncu-rep --page raw --csv > gpurun_out/r02_persist_full_raw.csv 2>/dev/null
GCBF_PERSISTENT=0 timeout 600 ncu --set full --clock-control none -k regex:edge_chain_kernel -s 30 -c 2 -o gpurun_out/r02_edge_chain_full -f \
   python bench.py --steps 1 --warmup 3 --T 16 --no-cpu-baseline --no-train > gpurun_out/r02_ncu_full_edge.log 2>&1
ncu -i gpurun_out/r02_edge_chain_full.ncu-rep --page raw --csv > gpurun_out/r02_edge_chain_full_raw.csv 2>/dev/null
python - <<'PY'
import json
for f in ("r02_bench_final","r02_bench_reference_arm"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); print(f, d["value"], d.get("e2e",{}).get("value"), d.get("gpu_launches"), (d.get("train_step") or {}).get("ms_per_minibatch"))
    except Exception as e: print(f, "ERR", e)
PY
ls -la gpurun_out | tail -15
