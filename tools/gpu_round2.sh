#!/bin/bash
# Round-2 GPU evidence (run under gpurun from the repo root):  bash tools/gpu_round2.sh <stage>...
set -u
mkdir -p gpurun_out
for stage in "$@"; do
case "$stage" in
  tests)    timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -12 ;;
  newtests) timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -25 ;;
  sanitize) bash tools/gpu_sanitize.sh 2>&1 | tee gpurun_out/sanitize_summary.txt ;;
  bench)    timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_full.json ;;
  benchq)   timeout 600 python bench.py --tuples 8000000 --steps 5 --warmup 3 --no-cpu --e2e-tuples 1000000 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "benchq rc=$?"; cut -c1-300 gpurun_out/bench_q.json ;;
  ref)      timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/bench_reference.json ;;
  multitests) timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -k "multi or ring or early or kats or cpp_host or queue or stream" 2>&1 | tee gpurun_out/pytest_gpu_multi.log | tail -15 ;;
  benchN)   N=${DTE_N:-2}; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --tuples ${DTE_TUPLES:-8000000} --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench$N rc=$?"; cut -c1-300 gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err ;;
  autotune) timeout 600 python tools/autotune_report.py 2>&1 | tee gpurun_out/autotune_report.txt ;;
  memmulti) timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 8 python -m pytest tests/test_gpu_multi.py -q -x -k "gpus000 or ring_combine or queue or refuses" > gpurun_out/sanitizer_memcheck_multi.log 2>&1; echo "memcheck multi rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_memcheck_multi.log | tr '\n' ' ')" ;;
  micro)    timeout 300 tools/pipe_microbench > gpurun_out/pipe_microbench.json 2> gpurun_out/pipe_microbench.err; echo "micro rc=$?"; cat gpurun_out/pipe_microbench.json ;;
  launches) B="--tuples 2000000 --steps 2 --warmup 1 --no-cpu --e2e-tuples 200000 --no-extras"
            timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py $B > gpurun_out/launches_bench.log 2>&1; echo "ncu launches rc=$?" ;;
  ncu)      B="--tuples 2000000 --steps 2 --warmup 1 --no-cpu --e2e-tuples 200000"
            timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dt_walk_tile -s 2 -c 1 -f -o gpurun_out/prof_walk python bench.py $B > gpurun_out/prof_walk.log 2>&1; echo "ncu full rc=$?"
            timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py $B > gpurun_out/launches_bench.log 2>&1; echo "ncu launches rc=$?" ;;
  *) echo "unknown stage $stage" ;;
esac
done
