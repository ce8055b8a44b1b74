#!/bin/bash
# DEIM evidence: (1) bench with the full-size Burgers DEIM (reduced POD rows), (2) ncu launch list of the
# reduced-rows bench command, (3) ncu --set full of the blocked-path kernels at 16M rows
mkdir -p gpurun_out
timeout 300 python bench.py --rows 1048576 --steps 2 --warmup 3 --no-e2e --no-cpu --no-rsvd --no-kat > gpurun_out/bench_deim.json 2> gpurun_out/bench_deim.err; echo "bench exit $?" >> gpurun_out/bench_deim.err
CMD="python bench.py --rows 2097152 --deim-rows 2097152 --steps 1 --warmup 1 --no-e2e --no-cpu --no-rsvd --no-kat"
timeout 200 $CMD > gpurun_out/plain.log 2> gpurun_out/plain.err &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
PB="python modelorderreduction.jl_b200/tools/panel_bench.py 16777216 256"
timeout 100 $PB > gpurun_out/plain3.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"deim_panel_kernel" -s 7 -c 1 -o gpurun_out/prof_deim_panel -f $PB > gpurun_out/ncu_panel.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"deim_step4_kernel" -s 20 -c 1 -o gpurun_out/prof_deim_step4 -f $PB > gpurun_out/ncu_step4.log 2>&1
tail -n 2 gpurun_out/bench_deim.err gpurun_out/ncu_launches.log gpurun_out/ncu_panel.log gpurun_out/ncu_step4.log
cat gpurun_out/bench_deim.json
