#!/bin/bash
# The one GPU-side job script (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash modelorderreduction.jl_b200/tools/gpu_job.sh tests smoke bench'
# Stages (any subset, run in the order given; every stage writes its log under gpurun_out/):
#   tests            python -m pytest tests -m gpu -x -q
#   smoke            __graft_entry__.smoke()
#   bench            python bench.py                      (full default run)
#   bench-quick      python bench.py --steps 2 --warmup 3 --no-kat --no-rsvd --e2e-steps 1
#   ref              python bench.py --impl reference --steps 2 --warmup 1
#   multi            torchrun over all visible GPUs: tests/multi_gpu_check.py, then bench.py --gpus N (quick)
#   ncu-launches     launch list (gpu__time_duration.sum) of TOOL (env, default tools/deim_once.py)
#   ncu-full         ncu --set full of kernels matching KERNEL (env regex) in TOOL, launch-skip SKIP, count COUNT
#   run              just run TOOL (env) with ARGS (env)
# Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
T=modelorderreduction.jl_b200/tools
TOOL=${TOOL:-$T/deim_once.py}
ARGS=${ARGS:-}
NG=$(python -c 'import torch; print(torch.cuda.device_count())' 2>/dev/null || echo 1)
for stage in "$@"; do
  case "$stage" in
    tests)
      timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.txt 2>&1
      echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt; tail -n 5 gpurun_out/pytest_gpu.txt ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
      echo "smoke exit $?" >> gpurun_out/smoke.txt; tail -n 3 gpurun_out/smoke.txt ;;
    bench)
      ( time timeout 1700 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2> gpurun_out/bench_time.txt
      echo "bench exit $?" >> gpurun_out/bench_full.err; tail -n 3 gpurun_out/bench_full.err gpurun_out/bench_time.txt; cat gpurun_out/bench_full.json ;;
    bench-quick)
      timeout 900 python bench.py --steps 2 --warmup 3 --no-kat --no-rsvd --e2e-steps 1 ${BENCH_ARGS:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
      echo "bench-quick exit $?" >> gpurun_out/bench_quick.err; tail -n 3 gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json ;;
    ref)
      timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
      echo "ref exit $?" >> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json ;;
    multi)
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NG" --master-addr 127.0.0.1 --master-port 29617 \
        tests/multi_gpu_check.py > gpurun_out/multi_check_n$NG.txt 2>&1
      echo "multi_gpu_check exit $?" >> gpurun_out/multi_check_n$NG.txt; tail -n 6 gpurun_out/multi_check_n$NG.txt
      MOR_B200_DEIM_EXCHANGE=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NG" --master-addr 127.0.0.1 \
        --master-port 29619 tests/multi_gpu_check.py > gpurun_out/multi_check_nccl_n$NG.txt 2>&1
      echo "multi_gpu_check (NCCL exchange) exit $?" >> gpurun_out/multi_check_nccl_n$NG.txt; tail -n 3 gpurun_out/multi_check_nccl_n$NG.txt
      timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NG" --master-addr 127.0.0.1 --master-port 29618 \
        bench.py --gpus "$NG" --steps 3 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err
      echo "bench N=$NG exit $?" >> gpurun_out/bench_n$NG.err; tail -n 3 gpurun_out/bench_n$NG.err; cat gpurun_out/bench_n$NG.json ;;
    ncu-launches)
      timeout 600 python $TOOL $ARGS > gpurun_out/tool_plain.txt 2>&1 && \
      timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c ${COUNT:-2000} --csv \
        --log-file gpurun_out/launches.csv python $TOOL $ARGS > gpurun_out/ncu_launches.log 2>&1
      echo "ncu-launches exit $?"; tail -n 2 gpurun_out/tool_plain.txt ;;
    ncu-full)
      timeout 600 python $TOOL $ARGS > gpurun_out/tool_plain.txt 2>&1 && \
      timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:${KERNEL:-.}" \
        --launch-skip ${SKIP:-0} --launch-count ${COUNT:-1} -f -o gpurun_out/${OUT:-full} python $TOOL $ARGS > gpurun_out/ncu_full.log 2>&1
      echo "ncu-full exit $?"; tail -n 3 gpurun_out/ncu_full.log ;;
    run)
      timeout ${RUN_TIMEOUT:-900} python $TOOL $ARGS > gpurun_out/${OUT:-run}.txt 2>&1
      echo "run exit $?" >> gpurun_out/${OUT:-run}.txt; tail -n ${TAIL:-40} gpurun_out/${OUT:-run}.txt ;;
    *) echo "unknown stage $stage" ;;
  esac
done
