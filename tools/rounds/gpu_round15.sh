#!/bin/bash
# accuracy forensics of the tcgen05 path at M = 1024 + the vectorised K1 planes builder
TAG=${1:-r01u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
timeout 300 python tools/tc_probe.py 1000000 1024 2 2>&1 | head -2 > $OUT/k1_probe.jsonl; cat $OUT/k1_probe.jsonl
timeout 900 python tests/probes/accum_probe.py 32768 1024 2 > $OUT/accum_probe.jsonl 2> $OUT/accum_probe.err; cat $OUT/accum_probe.jsonl; tail -3 $OUT/accum_probe.err
timeout 900 python tests/probes/parity_probe.py 32768,1024,2,0,0,0 32768,512,4,0,0,0 16384,256,4,0,0,0 > $OUT/parity_simt.jsonl 2> $OUT/parity_simt.err; cat $OUT/parity_simt.jsonl; tail -3 $OUT/parity_simt.err
