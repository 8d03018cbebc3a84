#!/bin/bash
TAG=${1:-r02v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests -m gpu -q -x --timeout 150 > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
python -c "
import json
d=json.load(open('$OUT/bench.json')); print(round(d['value']), round(d['ms_per_step'],1), d['roofline']['frac'], round(d['e2e']['value']), d['clocks']); print(d['kernels_ms'])
" || tail -5 $OUT/bench.err
