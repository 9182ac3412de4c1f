#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/mp_train_worker.py gpurun_out/r02_multi_res.json > gpurun_out/r02_multi_worker.log 2>&1
echo "exit $?" >> gpurun_out/r02_multi_worker.log
tail -25 gpurun_out/r02_multi_worker.log; cat gpurun_out/r02_multi_res.json 2>/dev/null
