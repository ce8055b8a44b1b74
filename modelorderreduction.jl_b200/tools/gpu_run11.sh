#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_deim_gpu.py tests/test_pod_gpu.py -q -m gpu -p no:cacheprovider -k "deim or fitzhugh" > gpurun_out/pytest_deim.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_deim.txt
timeout 200 python modelorderreduction.jl_b200/tools/deim_e2e.py > gpurun_out/deim_e2e.txt 2>&1; echo "exit $?" >> gpurun_out/deim_e2e.txt
tail -n 5 gpurun_out/pytest_deim.txt
cat gpurun_out/deim_e2e.txt
