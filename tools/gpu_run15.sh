#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dubins_diag.py > gpurun_out/r02_dubins_diag.log 2>&1; cat gpurun_out/r02_dubins_diag.log | tail -20
