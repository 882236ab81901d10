#!/bin/bash
mkdir -p gpurun_out
BARGS="--tuples 2000000 --steps 2 --warmup 1 --no-cpu --e2e-tuples 200000"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dt_walk_tile -s 2 -c 1 -f -o gpurun_out/prof_staged_v2b python bench.py $BARGS > gpurun_out/prof_staged_v2b.log 2>&1; echo rc=$?
timeout 600 python bench.py --tuples 8000000 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_v2b.json 2> gpurun_out/bench_v2b.err; cut -c1-200 gpurun_out/bench_v2b.json
