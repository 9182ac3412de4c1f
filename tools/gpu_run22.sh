#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py tests/test_gpu_gnn.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r02_pytest_22.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_22.log
timeout 600 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k "persistent" > gpurun_out/r02_pytest_22b.log 2>&1
echo "pytest rollout rc=$?"; tail -2 gpurun_out/r02_pytest_22b.log
timeout 600 python bench.py --train-only 2> gpurun_out/r02_train_only22.err | cut -c1-300
GCBF_TRAIN_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches22.csv python bench.py --train-only --T 8 > gpurun_out/r02_train_ncu22.log 2>&1
