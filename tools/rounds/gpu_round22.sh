#!/bin/bash
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for cfg in "1 0" "1 4" "1 2" "0 0"; do
  set -- $cfg
  echo "KCACHE=$1 KSEG=$2" | tee -a $OUT/parity_kcache.jsonl
  SVGP_SCALED_KCACHE=$1 SVGP_SCALED_KSEG=$2 SVGP_SCALED_KSEG2=0 timeout 300 python tests/probes/parity_probe.py 32768,1024,2 16384,256,4 2>&1 | grep -v Warn | tail -2 | tee -a $OUT/parity_kcache.jsonl
done
SVGP_SCALED_KSEG=4 SVGP_SCALED_KSEG2=0 timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for cfg in "1 0" "1 4" "1 2" "1 8"; do
  set -- $cfg
  SVGP_SCALED_KCACHE=$1 SVGP_SCALED_KSEG=$2 SVGP_SCALED_KSEG2=0 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_c$1_k$2.json 2> $OUT/bench_c$1_k$2.err
  python -c "
import json
d=json.load(open('$OUT/bench_c$1_k$2.json')); print('KCACHE $1 KSEG $2', round(d['value']), round(d['ms_per_step'],1), {k:v for k,v in list(d['kernels_ms'].items())[:3]})
" || tail -3 $OUT/bench_c$1_k$2.err
done
