#!/bin/bash
# usage: gpu_run_mbench_quick.sh N  (under gpurun --gpus N): bench at N ranks without the e2e / KAT / RSVD / CPU sections
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --no-kat --no-rsvd --no-cpu > gpurun_out/bench_quick_n$N.json 2> gpurun_out/bench_quick_n$N.err; echo "bench exit $?" >> gpurun_out/bench_quick_n$N.err
tail -n 3 gpurun_out/bench_quick_n$N.err; cat gpurun_out/bench_quick_n$N.json
