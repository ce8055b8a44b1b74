#!/bin/bash
# First contact with the GPU box: machine facts, FP64 calibration, DEIM parity + timing.
mkdir -p gpurun_out
{
  nvidia-smi; nproc; free -g; lscpu | head -20
  nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks.mem,power.draw,power.limit --format=csv
} > gpurun_out/probe.txt 2>&1
timeout 120 modelorderreduction.jl_b200/tools/ubench_fp64 > gpurun_out/ubench.txt 2>&1
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 300 python modelorderreduction.jl_b200/tools/deim_bench.py > gpurun_out/deim_bench.txt 2>&1
tail -5 gpurun_out/ubench.txt gpurun_out/pytest_gpu.txt gpurun_out/deim_bench.txt
