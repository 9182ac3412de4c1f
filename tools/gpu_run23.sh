#!/bin/bash
mkdir -p gpurun_out
for B in 32 64; do
GCBF_BENCH_TRAIN_GRAPHS=$B timeout 600 python bench.py --train-only 2> gpurun_out/r02_train_only23_$B.err | cut -c1-330
done
GCBF_BENCH_TRAIN_GRAPHS=32 GCBF_TRAIN_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches23_B32.csv python bench.py --train-only --T 8 > gpurun_out/r02_train_ncu23.log 2>&1
python tools/train_launch_summary.py gpurun_out/r02_train_launches23_B32.csv
