#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_gnn.py tests/test_gpu_fullsize.py tests/test_gpu_rollout.py tests/test_gpu_trainer.py tests/test_gpu_qp.py -x -q -m gpu > gpurun_out/r02_pytest_18.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_18.log
for mode in "16 16" "32 16" "16 32" "32 32"; do
  set -- $mode
  GCBF_TC_BK=$1 GCBF_TC_TN_ROWS=$2 timeout 600 python bench.py --train-only --T 64 2> gpurun_out/r02_train_only18_$1_$2.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)['train_step']; print('BK=$1 TN_ROWS=$2', d['ms_per_minibatch'], d['kernels_per_step'], d['edges_per_rank'])"
done
