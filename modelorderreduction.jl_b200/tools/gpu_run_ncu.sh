#!/bin/bash
# final-build ncu evidence: launch list of a reduced bench command + full captures of the hot kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pod_gpu.py -q -m gpu -p no:cacheprovider -k "abi_error or known_answer" > gpurun_out/pytest_new.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_new.txt
CMD="python bench.py --rows 2097152 --deim-rows 2097152 --steps 1 --warmup 1 --no-e2e --no-cpu --no-rsvd --no-kat"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2> gpurun_out/plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tn_kernel|jacobi_block_sweep|deim_update" -c 6 -o gpurun_out/prof_final -f $CMD > gpurun_out/ncu_final.log 2>&1
tail -n 3 gpurun_out/pytest_new.txt gpurun_out/ncu_launches.log gpurun_out/ncu_final.log
