#!/bin/bash
TAG=${1:-r02i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 500 python -m pytest tests -m gpu -q -x --timeout 150 > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
python -c "
import json
d=json.load(open('$OUT/bench.json')); print(round(d['value']), round(d['ms_per_step'],1), d['roofline']['frac'], d['roofline_k1'], d['e2e']['value'], d['cpu_baseline']['value']); print(d['kernels_ms'])
" || tail -5 $OUT/bench.err
SVGP_SCALED_KSEG=8 timeout 200 python tests/probes/parity_probe.py 32768,1024,2 65536,1024,2 2>/dev/null | tee $OUT/parity_kseg8.jsonl
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err; head -c 600 $OUT/bench_ref.json; tail -4 $OUT/bench_ref.err
