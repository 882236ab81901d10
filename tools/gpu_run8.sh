#!/bin/bash
mkdir -p gpurun_out
run() { # tune variant extra
  f=gpurun_out/run8_$(echo "$1" | tr ',=' '__')_v$2$4.json
  DTE_TUNE=$1 timeout 400 python bench.py --variant $2 --tuples 4000000 --steps 3 --warmup 3 --e2e-tuples 200000 $3 > $f 2> $f.err
  python -c "
import json
try:
    d=json.load(open('$f')); print('$1 v$2 $3 ->', round(d['value']/1e6,2), 'M/s', d['config']['kernel'], d['config']['tuples_per_cta'], 'frac', round(d['roofline']['frac'],3), d['parity_spot_check'])
except Exception as e: print('$1 v$2 fail', e)
"
}
run "pair=2,ilp=2,stages=2" 3
run "pair=2,ilp=2,stages=1" 3
run "pair=2,ilp=2,stages=3" 3
run "pair=2,ilp=4,stages=1" 3
