#!/bin/bash
# blocked DEIM + Burgers generator: new tests first, then the DEIM timing tool
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_deim_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_deim.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_deim.txt
timeout 200 python modelorderreduction.jl_b200/tools/deim_bench.py > gpurun_out/deim_bench2.txt 2>&1; echo "deim_bench exit $?" >> gpurun_out/deim_bench2.txt
tail -n 30 gpurun_out/pytest_deim.txt
cat gpurun_out/deim_bench2.txt
