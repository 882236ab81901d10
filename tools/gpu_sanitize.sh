#!/bin/bash
# compute-sanitizer over the walk kernel that SHIPS for the headline geometry (D = 12, F = 256, AUTO plan =
# dt_walk_tile<4,2,1,0,384>, phased ring refill, several tiles per CTA) and over the planner's alternatives.
# Logs (with the kernel name printed by the test) go to gpurun_out/sanitizer_<tool>_<plan>.log.
set -u
mkdir -p gpurun_out
T=tests/test_gpu_parity.py::test_shipped_plan_headline_geometry_multi_tile
for tool in racecheck synccheck memcheck; do
  for tune in "" "pair=4,stages=1" "pair=1,ilp=8,stages=1" "pair=2,ilp=2,stages=2"; do
    tag=$(echo "${tune:-auto}" | tr ',=' '__')
    log=gpurun_out/sanitizer_${tool}_${tag}.log
    extra=""
    [ "$tool" = racecheck ] && extra="--racecheck-report all"
    DTE_TUNE="$tune" timeout 900 compute-sanitizer --tool $tool $extra --error-exitcode 7 --print-limit 20 \
        python -m pytest "$T" -q -x -s > "$log" 2>&1
    echo "$tool [$tag] rc=$? $(grep -E 'shipped-plan|ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' "$log" | tr '\n' ' ' | cut -c1-400)"
  done
done
