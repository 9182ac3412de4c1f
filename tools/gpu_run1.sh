#!/bin/bash
# round-2 GPU pass 1: parity tests, bench line, sanitizers, ncu launch list + full capture of the dominant kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt 2>&1
( time timeout 1700 python -m pytest tests -m gpu -q --maxfail=20 -x --deselect tests/test_gpu_multi.py ) > gpurun_out/r02_pytest1.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_pytest1.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
for tool in memcheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_target.py DoubleIntegrator > gpurun_out/r02_sanitizer_${tool}_DI.log 2>&1
  echo "exit $?" >> gpurun_out/r02_sanitizer_${tool}_DI.log
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/r02_launches_T16.csv \
   python bench.py --steps 2 --warmup 1 --T 16 --no-cpu-baseline --no-train > gpurun_out/r02_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_prod -s 20 -c 2 -o gpurun_out/r02_prod_full -f \
   python bench.py --steps 1 --warmup 1 --T 16 --no-cpu-baseline --no-train > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_prod_full.ncu-rep --page raw --csv > gpurun_out/r02_prod_full_raw.csv 2>/dev/null
ls -la gpurun_out
