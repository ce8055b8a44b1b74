#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt
( time timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2> gpurun_out/bench_time.txt; echo "bench full exit $?" >> gpurun_out/bench_full.err
tail -n 4 gpurun_out/pytest_gpu.txt gpurun_out/smoke.txt gpurun_out/bench_full.err gpurun_out/bench_time.txt
cat gpurun_out/bench_full.json
