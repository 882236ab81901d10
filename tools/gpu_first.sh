#!/bin/bash
# first GPU contact: smoke, parity tests, quick bench sweep.  Every step has its own timeout.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -5 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/pytest_gpu.log
echo "== bench sweep"
for v in 2 3; do
  timeout 400 python bench.py --variant $v --tuples 4000000 --steps 3 --warmup 3 --no-cpu --e2e-tuples 500000 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err ; echo "v$v rc=$?"; cat gpurun_out/bench_v$v.json | cut -c1-400
done
for tune in "ilp=4,stages=2" "ilp=4,stages=3" "ilp=8,stages=2" "ilp=4,stages=4,warps=3" "ilp=4,stages=2,warps=4" "ilp=4,stages=2,warps=3"; do
  f=gpurun_out/bench_tune_$(echo $tune | tr ',=' '__').json
  DTE_TUNE=$tune timeout 400 python bench.py --variant 3 --tuples 4000000 --steps 3 --warmup 3 --no-cpu --e2e-tuples 500000 > $f 2> $f.err ; echo "$tune rc=$?"; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(d['value'], d['config']['kernel'], d['config']['tuples_per_cta'], d['roofline']['frac'])
except Exception as e: print('fail', e)
"
done
