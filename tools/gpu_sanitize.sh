#!/bin/bash
# compute-sanitizer over the walk kernel that SHIPS for the headline geometry (D = 12, F = 256, AUTO plan =
# dt_walk_tile<4,2,1,0,384>, phased ring refill, CTAs with two tiles) and over the planner's alternatives.
# Logs (with the kernel name printed by the test) go to gpurun_out/sanitizer_<tool>_<plan>.log.
#   bash tools/gpu_sanitize.sh                 memcheck + synccheck over four launch plans, racecheck once (AUTO)
# racecheck does not model the mbarrier / cp.async.bulk (async-proxy) hand-off: tools/racecheck_canonical.cu is the
# textbook producer/consumer ring with exactly that synchronisation, run under the same tool for comparison.
set -u
mkdir -p gpurun_out
export DTE_TEST_SANITIZER=1
T=tests/test_gpu_parity.py::test_shipped_plan_headline_geometry_multi_tile
run() { # tool tune extra-args
  tag=$(echo "${2:-auto}" | tr ',=' '__')
  log=gpurun_out/sanitizer_$1_${tag}.log
  DTE_TUNE="$2" timeout 900 compute-sanitizer --tool $1 $3 --error-exitcode 7 --print-limit 8 --show-backtrace no \
      python -m pytest "$T" -q -x -s > "$log" 2>&1
  echo "$1 [$tag] rc=$? $(grep -E 'shipped-plan|ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' "$log" | tr '\n' ' ' | cut -c1-400)"
}
for tool in memcheck synccheck; do
  for tune in "" "pair=4,stages=1" "pair=1,ilp=8,stages=1" "pair=2,ilp=2,stages=2"; do run $tool "$tune" ""; done
done
timeout 300 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 8 --show-backtrace no tools/racecheck_canonical > gpurun_out/sanitizer_racecheck_canonical.log 2>&1
echo "racecheck [canonical textbook pipeline] rc=$? $(grep -E 'canonical|RACECHECK SUMMARY' gpurun_out/sanitizer_racecheck_canonical.log | tr '\n' ' ' | cut -c1-300)"
run racecheck "" "--racecheck-report all"
