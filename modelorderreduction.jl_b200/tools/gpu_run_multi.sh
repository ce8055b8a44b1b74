#!/bin/bash
# usage: gpu_run_multi.sh N   (under gpurun --gpus N): sharded parity check + bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/mgpu_${N}_info.txt; free -g >> gpurun_out/mgpu_${N}_info.txt; nproc >> gpurun_out/mgpu_${N}_info.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/mgpu_check_n$N.txt 2>&1; echo "check exit $?" >> gpurun_out/mgpu_check_n$N.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?" >> gpurun_out/bench_n$N.err
grep -v "^W1\|^\[W\|^\*\*\*\|OMP_NUM" gpurun_out/mgpu_check_n$N.txt | tail -n 8; tail -n 3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
