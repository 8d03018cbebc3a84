#!/bin/bash
TAG=${1:-r02o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 120 python -m pytest tests/test_gpu_tc_engine.py -m gpu -q -x --timeout 60 > $OUT/pytest_tc.log 2>&1; tail -4 $OUT/pytest_tc.log
timeout 120 python tools/tc_probe.py 1000000 1024 64 syrk 2>&1 | grep '"chunk": 512\|"chunk": 2048\|rror' | tee $OUT/syrk_probe.jsonl
SVGP_SYRK_CLUSTER=1 timeout 120 python tools/tc_probe.py 1000000 1024 64 syrk 2>&1 | grep '"chunk": 512\|rror' | tee -a $OUT/syrk_probe.jsonl
