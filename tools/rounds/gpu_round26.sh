#!/bin/bash
TAG=${1:-r02g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x --timeout 120 -k "graph" > $OUT/pytest_graph.log 2>&1; tail -15 $OUT/pytest_graph.log
timeout 300 python tools/small_configs_timing.py > $OUT/small_configs.jsonl 2> $OUT/small_configs.err; cat $OUT/small_configs.jsonl; tail -5 $OUT/small_configs.err
