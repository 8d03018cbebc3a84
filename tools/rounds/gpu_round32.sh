#!/bin/bash
TAG=${1:-r02q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for sc in 16384 20480 24576 28672 32768 40960 49152; do
  echo "SVGP_SYRK_SC=$sc" | tee -a $OUT/syrk_sc.jsonl
  SVGP_SYRK_SC=$sc timeout 100 python tools/tc_probe.py 1000000 1024 64 syrk 2>&1 | grep '"chunk": 512\|rror' | tee -a $OUT/syrk_sc.jsonl
done
