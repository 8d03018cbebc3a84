#!/bin/bash
# accuracy / time trade of the chain lengths: SCALED k-segments per matrix family, SYRK chunk rows
TAG=${1:-r01w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for cfg in "4 4 512" "0 4 512" "4 0 512" "2 2 512" "1 1 512" "4 4 128" "2 2 128" "1 1 128" "1 1 64"; do
  set -- $cfg
  echo "KSEG=$1 KSEG2=$2 chunk=$3" | tee -a $OUT/parity_matrix.jsonl
  SVGP_SCALED_KSEG=$1 SVGP_SCALED_KSEG2=$2 timeout 300 python tests/probes/parity_probe.py 32768,1024,2,$3 2>/dev/null | tee -a $OUT/parity_matrix.jsonl
done
timeout 600 python tests/probes/parity_fullsize.py 262144 1024 2 > $OUT/parity_fullsize.jsonl 2> $OUT/parity_fullsize.err; cat $OUT/parity_fullsize.jsonl; tail -3 $OUT/parity_fullsize.err
timeout 300 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/syrk_chunk_sweep.jsonl 2>&1; cat $OUT/syrk_chunk_sweep.jsonl
for cfg in "0 0" "0 4" "0 2" "0 1" "4 4" "4 2" "2 2"; do
  set -- $cfg
  SVGP_SCALED_KSEG=$1 SVGP_SCALED_KSEG2=$2 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_k$1_$2.json 2> $OUT/bench_k$1_$2.err
  python -c "
import json
d=json.load(open('$OUT/bench_k$1_$2.json')); print('KSEG $1 KSEG2 $2', round(d['value']), round(d['ms_per_step'],1), {k:v for k,v in list(d['kernels_ms'].items())[:3]})
"
done
