#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full exit $?" >> gpurun_out/bench_full.err
CMD="python modelorderreduction.jl_b200/tools/deim_once.py 4194304 256"
$CMD > gpurun_out/deim_once.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:deim_step_kernel -s 200 -c 2 -o gpurun_out/prof_deim -f $CMD > gpurun_out/ncu_deim.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.txt gpurun_out/bench_full.err gpurun_out/deim_once.log gpurun_out/ncu_deim.log
cat gpurun_out/bench_full.json
