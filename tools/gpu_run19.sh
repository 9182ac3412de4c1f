#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py tests/test_gpu_gnn.py tests/test_gpu_fullsize.py tests/test_gpu_qp.py -x -q -m gpu > gpurun_out/r02_pytest_19.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r02_pytest_19.log
timeout 600 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k "persistent" > gpurun_out/r02_pytest_19b.log 2>&1
echo "pytest rollout rc=$?"; tail -3 gpurun_out/r02_pytest_19b.log
for fold in 1; do
  GCBF_TRAIN_FOLD=$fold GCBF_TC_BK=32 timeout 600 python bench.py --train-only --T 64 2> gpurun_out/r02_train_only19_$fold.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)['train_step']; print('FOLD=$fold', d['ms_per_minibatch'], d['kernels_per_step'], d['edges_per_rank'])"
done
