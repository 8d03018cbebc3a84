#!/bin/bash
TAG=${1:-r02s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 200 python tests/probes/parity_probe.py 32768,1024,2 65536,1024,2 32768,512,4 2>/dev/null | tee $OUT/parity_probe.jsonl
timeout 300 python tests/probes/parity_fullsize.py 524288 1024 2 2>/dev/null | tee $OUT/parity_fullsize.jsonl
timeout 400 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
python -c "
import json
d=json.load(open('$OUT/bench.json')); print(round(d['value']), round(d['ms_per_step'],1), d['roofline']['frac'], d['e2e']['value']); print(d['kernels_ms'])
" || tail -5 $OUT/bench.err
timeout 300 python -m pytest tests -m gpu -q -x --timeout 150 > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
