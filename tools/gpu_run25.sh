#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
GCBF_TC_BN128=$v timeout 600 python bench.py --train-only 2> gpurun_out/r02_train_only25_$v.err | cut -c1-120
done
