#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r02_8gpu_devices.txt
run() { # name, extra args
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 8 --steps 5 --warmup 3 ${@:3} > gpurun_out/$1.json 2> gpurun_out/$1.err
  echo "$1 exit $?"
}
run r02_bench_8gpu_config3 29521
run r02_bench_8gpu_config4 29522 --config 4
run r02_bench_8gpu_config5 29523 --config 5 --steps 3
( timeout 500 python -m pytest tests/test_gpu_multi.py -q ) > gpurun_out/r02_pytest14_multi.log 2>&1; tail -3 gpurun_out/r02_pytest14_multi.log
python - <<'PY'
import json
for f in ("r02_bench_8gpu_config3","r02_bench_8gpu_config4","r02_bench_8gpu_config5"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); t=d.get("train_step") or {}
        print(f, d["n_gpus"], d["value"], d["config"]["us_per_env_step"], t.get("ms_per_minibatch"), t.get("graphs_per_rank"))
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
