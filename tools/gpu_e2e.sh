#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
for c in 16384 32768 65536 131072 262144 524288; do
  DTE_TUNE=chunk=$c timeout 300 python bench.py --tuples 4000000 --steps 3 --warmup 3 --no-cpu --e2e-tuples 4000000 > gpurun_out/e2e_$c.json 2> gpurun_out/e2e_$c.err
  python -c "
import json
d=json.load(open('gpurun_out/e2e_$c.json')); print('chunk $c tuples ->', round(d['e2e']['value']/1e6,2), 'M tuples/s e2e', round(d['e2e']['value']*1024/1e9,1),'GB/s h2d')"
done
