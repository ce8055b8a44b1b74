#!/bin/bash
mkdir -p gpurun_out
for MODE in 3; do
MOR_B200_GEMM_TN_STREAMK=$MODE timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_pod_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_sk$MODE.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sk$MODE.txt
MOR_B200_GEMM_TN_STREAMK=$MODE timeout 900 python bench.py --no-cpu --no-e2e --no-rsvd --no-kat --steps 3 --warmup 2 > gpurun_out/bench_sk$MODE.json 2> gpurun_out/bench_sk$MODE.err; echo "bench exit $?" >> gpurun_out/bench_sk$MODE.err
tail -n 3 gpurun_out/pytest_sk$MODE.txt gpurun_out/bench_sk$MODE.err
done
