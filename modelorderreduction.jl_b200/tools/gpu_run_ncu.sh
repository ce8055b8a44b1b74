#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 1500 python bench.py --no-cpu > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full exit $?" >> gpurun_out/bench_full.err
CMD="python bench.py --rows 2097152 --deim-rows 2097152 --steps 1 --warmup 1 --no-e2e --no-cpu --no-rsvd --no-kat"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2> gpurun_out/plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tn_sk_kernel" -c 2 -o gpurun_out/prof_final -f $CMD > gpurun_out/ncu_final.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.txt gpurun_out/bench_full.err gpurun_out/ncu_launches.log gpurun_out/ncu_final.log
