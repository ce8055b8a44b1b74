#!/bin/bash
# usage: gpu_run_multi.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/mgpu_${N}_info.txt; free -g >> gpurun_out/mgpu_${N}_info.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_mgpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mgpu.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?" >> gpurun_out/bench_n$N.err
tail -n 5 gpurun_out/pytest_mgpu.txt; tail -n 5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
