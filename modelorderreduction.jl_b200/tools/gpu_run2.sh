#!/bin/bash
# tests + smoke + reduced bench + full bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt
timeout 600 python bench.py --rows 1048576 --deim-rows 1048576 --steps 3 --warmup 2 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench small exit $?" >> gpurun_out/bench_small.err
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full exit $?" >> gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -3 gpurun_out/pytest_gpu.txt gpurun_out/smoke.txt gpurun_out/bench_small.err gpurun_out/bench_full.err
cat gpurun_out/bench_small.json gpurun_out/bench_full.json gpurun_out/bench_ref.json
