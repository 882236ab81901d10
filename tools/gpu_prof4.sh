#!/bin/bash
mkdir -p gpurun_out
BARGS="--tuples 2000000 --steps 2 --warmup 1 --no-cpu --e2e-tuples 200000"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dt_walk_tile -s 2 -c 1 -f -o gpurun_out/prof_pair_v4 python bench.py $BARGS > gpurun_out/prof_pair_v4.log 2>&1; echo rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v4.csv python bench.py $BARGS > gpurun_out/launches_v4_bench.log 2>&1; echo rc=$?
timeout 900 python bench.py > gpurun_out/bench_full_v4.json 2> gpurun_out/bench_full_v4.err; echo rc=$?; cut -c1-220 gpurun_out/bench_full_v4.json
