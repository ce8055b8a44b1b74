#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_deim_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_deim.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_deim.txt
timeout 200 python modelorderreduction.jl_b200/tools/panel_bench.py 16777216 256 > gpurun_out/panel_bench.txt 2>&1; echo "exit $?" >> gpurun_out/panel_bench.txt
tail -n 5 gpurun_out/pytest_deim.txt
cat gpurun_out/panel_bench.txt
