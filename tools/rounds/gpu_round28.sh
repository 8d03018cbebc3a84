#!/bin/bash
TAG=${1:-r02j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x --timeout 150 -k "ragged" > $OUT/pytest_ragged.log 2>&1; tail -4 $OUT/pytest_ragged.log
for m in 256 512 2048; do
  timeout 400 python bench.py --m $m --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_m$m.json 2> $OUT/bench_m$m.err
  python -c "
import json
d=json.load(open('$OUT/bench_m$m.json')); print('M=$m', round(d['value']), 'dp/s', round(d['ms_per_step'],1), 'ms', {k:round(v) for k,v in list(d['kernels_ms'].items())[:6]})
" || tail -3 $OUT/bench_m$m.err
done
