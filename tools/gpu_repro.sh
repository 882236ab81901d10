#!/bin/bash
# Reproduces the round-1 GPU evidence on a B200 box (run from the repo root, e.g. under gpurun):
#   bash tools/gpu_repro.sh tests      GPU parity tests through the C ABI
#   bash tools/gpu_repro.sh sweep      launch-plan sweep (trees/warp x warps/group x ring stages), 4 M tuples
#   bash tools/gpu_repro.sh profile    ncu --set full of the walk kernel + launch list + compute-sanitizer
#   bash tools/gpu_repro.sh bench      the full default bench line and the reference (CPU oracle) arm
# Everything is written under gpurun_out/; every step has its own timeout.
set -u
mkdir -p gpurun_out
BARGS="--tuples 2000000 --steps 2 --warmup 1 --no-cpu --e2e-tuples 200000"
run_plan() { # DTE_TUNE string, extra bench args, tag
  f=gpurun_out/sweep_$(echo "$1" | tr ',=' '__')$3.json
  DTE_TUNE=$1 timeout 400 python bench.py --variant 3 --tuples 4000000 --steps 3 --warmup 3 --no-cpu --e2e-tuples 200000 $2 > $f 2> $f.err
  python - "$f" "$1 $2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[2], "->", round(d["value"] / 1e6, 2), "M tuples/s,", d["config"]["tuples_per_cta"], "tuples/CTA, frac", round(d["roofline"]["frac"], 3))
except Exception as e:
    print(sys.argv[2], "failed:", e)
PY
}
case "${1:-tests}" in
  tests)
    timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -5 ;;
  sweep)
    for t in "pair=2,ilp=4,stages=1" "pair=2,ilp=4,stages=1,warps=8" "pair=4,stages=1" "pair=2,ilp=2,stages=2" \
             "pair=1,ilp=8,stages=1" "pair=1,ilp=8,stages=2" "pair=1,ilp=4,stages=2" "pair=1,ilp=4,stages=1,warps=6"; do run_plan "$t" "" ""; done
    run_plan "" "--trees 512 --depth 8 --features 128" _cfg2
    run_plan "" "--trees 1024 --depth 10 --features 256" _cfg4shard ;;
  profile)
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dt_walk_tile -s 2 -c 1 -f -o gpurun_out/prof_walk python bench.py $BARGS > gpurun_out/prof_walk.log 2>&1; echo "ncu full rc=$?"
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py $BARGS > gpurun_out/launches_bench.log 2>&1; echo "ncu launches rc=$?"
    timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "kats or cfg1 or tuple_count or register_and_line" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
    timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "cfg1 or kats" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" ;;
  bench)
    timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_full.json
    timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/bench_reference.json ;;
  *) echo "usage: $0 tests|sweep|profile|bench"; exit 2 ;;
esac
