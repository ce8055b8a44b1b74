#!/bin/bash
mkdir -p gpurun_out
timeout 200 python modelorderreduction.jl_b200/tools/panel_bench.py 16777216 256 2568 25610 25616 > gpurun_out/panel_sweep.txt 2>&1; echo "exit $?" >> gpurun_out/panel_sweep.txt
cat gpurun_out/panel_sweep.txt
