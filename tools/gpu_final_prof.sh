#!/bin/bash
# final evidence for the round: sanitizer, ncu on the shipped kernel, launch list
mkdir -p gpurun_out
BARGS="--tuples 2000000 --steps 2 --warmup 1 --no-cpu --e2e-tuples 200000"
echo "== compute-sanitizer memcheck (KATs + small random cases through the C ABI)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "kats or cfg1 or tuple_count or register_and_line" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
echo "== compute-sanitizer racecheck (shared-memory ring + tile)"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "cfg1 or kats" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dt_walk_tile -s 2 -c 1 -f -o gpurun_out/prof_staged_v3 python bench.py $BARGS > gpurun_out/prof_staged_v3.log 2>&1; echo rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v3.csv python bench.py $BARGS > gpurun_out/launches_v3_bench.log 2>&1; echo rc=$?
