#!/bin/bash
TAG=${1:-r01r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
bash tools/rounds/gpu_round11.sh $TAG
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench.json | head -c 3000; tail -3 $OUT/bench.err
