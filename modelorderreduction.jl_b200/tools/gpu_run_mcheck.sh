#!/bin/bash
# usage: gpu_run_mcheck.sh N   (under gpurun --gpus N): sharded parity check only
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/mgpu_check_n$N.txt 2>&1; echo "check exit $?" >> gpurun_out/mgpu_check_n$N.txt
grep -v "^W1\|^\[W\|^\*\*\*\|OMP_NUM" gpurun_out/mgpu_check_n$N.txt | tail -n 12
