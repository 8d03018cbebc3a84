#!/bin/bash
TAG=${1:-r02p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for env in "A=1" "SVGP_TC_BK=32" "SVGP_TC_DEBUG=1" "SVGP_SYRK_SC=24576" "SVGP_SYRK_SC=6144"; do
  echo "$env" | tee -a $OUT/syrk_knobs.jsonl
  env $env timeout 100 python tools/tc_probe.py 1000000 1024 64 syrk 2>&1 | grep '"chunk": 512\|rror' | tee -a $OUT/syrk_knobs.jsonl
done
