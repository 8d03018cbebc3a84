#!/bin/bash
TAG=${1:-r02t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for mb in 16 32 64 96 128 256; do
  echo "SVGP_QUAD_L2MB=$mb" | tee -a $OUT/quad_l2.jsonl
  SVGP_QUAD_L2MB=$mb timeout 100 python tools/tc_probe.py 1000000 1024 64 quad 2>&1 | grep 'rowquad_tc_tri\|rror' | tee -a $OUT/quad_l2.jsonl
done
