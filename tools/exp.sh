run() { f=gpurun_out/exp_$(echo "$1" | tr ',=' '__').json
  DTE_TUNE=$1 timeout 400 python bench.py --tuples 8000000 --steps 3 --warmup 3 --no-cpu --e2e-tuples 200000 $2 > $f 2> $f.err
  python -c "
import json
try:
    d=json.load(open('$f')); print('$1 $2 ->', round(d['value']/1e6,2), 'M/s', d['config']['tuples_per_cta'])
except Exception as e: print('$1 fail', e)"; }
mkdir -p gpurun_out
run "phased=1"; run "phased=2"; run "phased=3"; run "phased=4"
run "phased=1" "--trees 1024 --depth 11 --features 256"; run "phased=0" "--trees 1024 --depth 11 --features 256"
