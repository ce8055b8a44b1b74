#!/bin/bash
mkdir -p gpurun_out
for m in 2097152 4194304; do
timeout 100 python modelorderreduction.jl_b200/tools/panel_bench.py $m 256 >> gpurun_out/panel_small.txt 2>&1
done
cat gpurun_out/panel_small.txt
