#!/bin/bash
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for b in 0 0.2 0.27 0.34; do
  echo "SVGP_SCALED_BIAS=$b" | tee -a $OUT/parity_bias.jsonl
  SVGP_SCALED_BIAS=$b timeout 600 python tests/probes/parity_probe.py 32768,1024,2 65536,1024,2 32768,512,4 2>/dev/null | tee -a $OUT/parity_bias.jsonl
done
timeout 600 python tests/probes/parity_probe.py 2304,256,4 16384,256,4 32768,2048,2 2>/dev/null | tee -a $OUT/parity_bias.jsonl
timeout 600 python tests/probes/parity_fullsize.py 524288 1024 2 > $OUT/parity_fullsize.jsonl 2> $OUT/parity_fullsize.err; cat $OUT/parity_fullsize.jsonl
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
