#!/bin/bash
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for f in 64 16 8 4 2; do
  echo "SVGP_SYRK_FLUSH=$f" | tee -a $OUT/parity_flush.jsonl
  SVGP_SYRK_FLUSH=$f timeout 300 python tests/probes/parity_probe.py 32768,1024,2 65536,1024,2,256 2>/dev/null | tee -a $OUT/parity_flush.jsonl
  SVGP_SYRK_FLUSH=$f timeout 200 python tools/tc_probe.py 1000000 1024 64 syrk 2>/dev/null | grep '"chunk": 512\|"chunk": 256' | tee -a $OUT/parity_flush.jsonl
done
