#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02_pytest_multi26.log 2>&1
echo "multi rc=$?"; tail -3 gpurun_out/r02_pytest_multi26.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu_final.json 2> gpurun_out/r02_bench_2gpu_final.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_2gpu_final.json") if l.startswith("{")][-1])
print(d["value"], d["e2e"]["value"], d["train_step"]["ms_per_minibatch"], d["train_step"]["kernels_per_step"])
PY
