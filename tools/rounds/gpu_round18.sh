#!/bin/bash
TAG=${1:-r01x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python tests/probes/ablate_probe.py 32768 1024 2 > $OUT/ablate.jsonl 2> $OUT/ablate.err; cat $OUT/ablate.jsonl; tail -3 $OUT/ablate.err
for cfg in "0 0" "4 4" "2 4"; do
  set -- $cfg
  SVGP_SCALED_KSEG=$1 SVGP_SCALED_KSEG2=$2 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_k$1_$2.json 2> $OUT/bench_k$1_$2.err
  python -c "
import json
d=json.load(open('$OUT/bench_k$1_$2.json')); print('KSEG $1 KSEG2 $2', round(d['value']), round(d['ms_per_step'],1), {k:v for k,v in list(d['kernels_ms'].items())[:3]})
"
done
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
