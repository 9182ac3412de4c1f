#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_multi.py ) > gpurun_out/r02_pytest2.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_pytest2.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke2.log 2>&1
tail -5 gpurun_out/r02_pytest2.log
