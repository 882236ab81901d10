#!/bin/bash
mkdir -p gpurun_out
run() { # tune variant
  f=gpurun_out/sweep2_$(echo "$1" | tr ',=' '__')_v$2.json
  DTE_TUNE=$1 timeout 400 python bench.py --variant $2 --tuples 4000000 --steps 3 --warmup 3 --no-cpu --e2e-tuples 200000 > $f 2> $f.err
  python -c "
import json
try:
    d=json.load(open('$f')); print('$1 v$2 ->', round(d['value']/1e6,2), 'M/s', d['config']['kernel'], d['config']['tuples_per_cta'], d['parity_spot_check'])
except Exception as e: print('$1 v$2 fail', e)
"
}
for t in "ilp=8,stages=1,warps=5" "ilp=8,stages=1,warps=4" "ilp=4,stages=1,warps=6" "ilp=4,stages=1,warps=5" "ilp=4,stages=2,warps=5"; do run $t 3; done
