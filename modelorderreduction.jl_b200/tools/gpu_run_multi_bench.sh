#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?" >> gpurun_out/bench_n$N.err
tail -n 5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
