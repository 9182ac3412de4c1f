#!/bin/bash
mkdir -p gpurun_out
run() { # name, port, extra args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 8 --steps 5 --warmup 3 ${@:3} > gpurun_out/$1.json 2> gpurun_out/$1.err
  echo "$1 exit $?"
}
run r02_bench_8gpu_config3_final 29531
run r02_bench_8gpu_config4_final 29532 --config 4
run r02_bench_8gpu_config5_final 29533 --config 5 --steps 3
python - <<'PY'
import json
for f in ("r02_bench_8gpu_config3_final","r02_bench_8gpu_config4_final","r02_bench_8gpu_config5_final"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); t=d.get("train_step") or {}
        print(f, d["n_gpus"], d["value"], d["config"]["us_per_env_step"], t.get("ms_per_minibatch"), t.get("graphs_per_rank"))
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
