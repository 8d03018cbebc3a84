#!/bin/bash
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -m gpu -q -x --timeout 120 > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 300 python tests/probes/parity_probe.py 32768,1024,2 65536,1024,2 32768,512,4 32768,2048,2 30000,1000,3 2>/dev/null | tee $OUT/parity_probe.jsonl
timeout 300 python tests/probes/parity_fullsize.py 524288 1024 2 2>/dev/null | tee $OUT/parity_fullsize.jsonl
timeout 400 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench.json | head -c 3600; tail -2 $OUT/bench.err
