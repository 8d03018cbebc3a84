#!/bin/bash
# DMMA DGEMM + prediction / Titsias entries: full GPU suite, DGEMM A/B, bench
TAG=${1:-r01l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
timeout 600 python tools/tc_probe.py 262144 1024 16 > $OUT/tc_probe.jsonl 2> $OUT/tc_probe.err
SVGP_DGEMM=simt timeout 600 python tools/tc_probe.py 262144 1024 16 > $OUT/tc_probe_simt.jsonl 2> $OUT/tc_probe_simt.err
grep "bmm64\|chol\|trinv" $OUT/tc_probe.jsonl $OUT/tc_probe_simt.jsonl
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench.json | head -c 3500; tail -3 $OUT/bench.err
