#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -5 gpurun_out/pytest_gpu.log
run() { # tune variant extra
  f=gpurun_out/run6_$(echo "$1" | tr ',=' '__')_v$2$4.json
  DTE_TUNE=$1 timeout 400 python bench.py --variant $2 --tuples 4000000 --steps 3 --warmup 3 --no-cpu --e2e-tuples 200000 $3 > $f 2> $f.err
  python -c "
import json
try:
    d=json.load(open('$f')); print('$1 v$2 $3 ->', round(d['value']/1e6,2), 'M/s', d['config']['kernel'], d['config']['tuples_per_cta'], 'frac', round(d['roofline']['frac'],3), d['parity_spot_check'])
except Exception as e: print('$1 v$2 fail', e)
"
}
run "ilp=4,pair=2,stages=1" 3
run "ilp=4,pair=2,stages=1,warps=8" 3
run "ilp=4,pair=2,stages=1,warps=6" 3
run "ilp=8,pair=1,stages=1" 3
run "ilp=4,pair=1,stages=1,warps=6" 3
run "" 0 "--trees 512 --depth 8 --features 128" _cfg2
run "ilp=8,pair=1,stages=1" 0 "--trees 512 --depth 8 --features 128" _cfg2
run "" 0 "--trees 1024 --depth 10 --features 256" _cfg4shard
