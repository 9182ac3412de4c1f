#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_2gpu_devices.txt
( timeout 900 python -m pytest tests/test_gpu_multi.py -q ) > gpurun_out/r02_pytest12_multi.log 2>&1; echo "exit $?" >> gpurun_out/r02_pytest12_multi.log; tail -5 gpurun_out/r02_pytest12_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_bench_2gpu.json") if l.startswith("{")][-1]); print(d["n_gpus"], d["value"], d["config"]["us_per_env_step"], d["train_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_bench_2gpu.err").read()[-1500:])
PY
