#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_rollout.py -q -x -k "persistent" ) > gpurun_out/r02_pytest3a.log 2>&1
echo "exit $?" >> gpurun_out/r02_pytest3a.log
tail -30 gpurun_out/r02_pytest3a.log
( timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py tests/test_gpu_rollout.py -q --maxfail=10 ) > gpurun_out/r02_pytest3b.log 2>&1
echo "exit $?" >> gpurun_out/r02_pytest3b.log
tail -5 gpurun_out/r02_pytest3b.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err
GCBF_PERSISTENT=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r02_bench3_5launch.json 2> gpurun_out/r02_bench3_5launch.err
head -c 400 gpurun_out/r02_bench3.json; tail -3 gpurun_out/r02_bench3.err
