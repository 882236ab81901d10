#!/bin/bash
# ncu evidence for the walk kernel + the full default bench line
mkdir -p gpurun_out
BARGS="--tuples 2000000 --steps 2 --warmup 1 --no-cpu --e2e-tuples 200000"
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py $BARGS > gpurun_out/launches_bench.log 2>&1; echo rc=$?
echo "== full set, staged kernel (launch #3 = first launch on the 2M-tuple batch)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dt_walk_tile -s 2 -c 1 -f -o gpurun_out/prof_staged python bench.py $BARGS > gpurun_out/prof_staged.log 2>&1; echo rc=$?
echo "== full set, tile (global top) kernel"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dt_walk_tile -s 2 -c 1 -f -o gpurun_out/prof_tile python bench.py --variant 2 $BARGS > gpurun_out/prof_tile.log 2>&1; echo rc=$?
echo "== full default bench"
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?; cat gpurun_out/bench_full.json | cut -c1-600
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo rc=$?; cat gpurun_out/bench_reference.json | cut -c1-400
nproc; lscpu | head -20 > gpurun_out/lscpu.txt
