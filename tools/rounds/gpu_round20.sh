#!/bin/bash
TAG=${1:-r01z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python tests/probes/accum_probe.py 32768 1024 2 2>/dev/null | grep -v "syrk\[" > $OUT/accum_scaled.jsonl; cat $OUT/accum_scaled.jsonl
timeout 600 python tests/probes/accum_probe.py 65536 1024 4 2>/dev/null | grep -v "syrk\[" > $OUT/accum_scaled_b.jsonl; cat $OUT/accum_scaled_b.jsonl
timeout 600 python tests/probes/accum_probe.py 32768 512 4 2>/dev/null | grep -v "syrk\[" > $OUT/accum_scaled_c.jsonl; cat $OUT/accum_scaled_c.jsonl
timeout 900 python tests/probes/parity_probe.py 2304,256,4 32768,512,4 32768,1024,2 65536,1024,2 32768,2048,2 > $OUT/parity_probe.jsonl 2> $OUT/parity_probe.err; cat $OUT/parity_probe.jsonl
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
