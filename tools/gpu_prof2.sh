#!/bin/bash
mkdir -p gpurun_out
BARGS="--tuples 2000000 --steps 2 --warmup 1 --no-cpu --e2e-tuples 200000"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dt_walk_tile -s 2 -c 1 -f -o gpurun_out/prof_staged_v2 python bench.py $BARGS > gpurun_out/prof_staged_v2.log 2>&1; echo rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v2.csv python bench.py $BARGS > gpurun_out/launches_v2_bench.log 2>&1; echo rc=$?
