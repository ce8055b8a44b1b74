#!/bin/bash
mkdir -p gpurun_out
MOR_B200_GEMM_TN_STREAMK=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_pod_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_sk1.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sk1.txt
MOR_B200_GEMM_TN_STREAMK=1 timeout 900 python bench.py --no-cpu --no-e2e --no-rsvd > gpurun_out/bench_sk1.json 2> gpurun_out/bench_sk1.err; echo "bench exit $?" >> gpurun_out/bench_sk1.err
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 1500 python bench.py --no-cpu > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full exit $?" >> gpurun_out/bench_full.err
tail -n 3 gpurun_out/pytest_sk1.txt gpurun_out/bench_sk1.err gpurun_out/pytest_gpu.txt gpurun_out/bench_full.err
