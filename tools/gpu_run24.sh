#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py tests/test_gpu_gnn.py -x -q -m gpu > gpurun_out/r02_pytest_24.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_24.log
for B in 256 32; do
GCBF_BENCH_TRAIN_GRAPHS=$B timeout 600 python bench.py --train-only 2> gpurun_out/r02_train_only24_$B.err | cut -c1-250
done
GCBF_BENCH_TRAIN_GRAPHS=32 GCBF_TRAIN_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches24_B32.csv python bench.py --train-only --T 8 > gpurun_out/r02_train_ncu24.log 2>&1
python tools/train_launch_summary.py gpurun_out/r02_train_launches24_B32.csv | head -14
