#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/persist_diag.py > gpurun_out/r02_persist_diag.log 2>&1
E=8 timeout 300 python tools/persist_diag.py > gpurun_out/r02_persist_diag_E8.log 2>&1
cat gpurun_out/r02_persist_diag.log
