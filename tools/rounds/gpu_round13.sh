#!/bin/bash
# sanity round after a container restore: gpu tests, smoke, bench (both arms)
TAG=${1:-r01s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
( time timeout 1200 python -m pytest tests -m gpu -q --durations=15 ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -30 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench.json | head -c 3500; tail -3 $OUT/bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
cat $OUT/bench_ref.json | head -c 1500; tail -4 $OUT/bench_ref.err
nvidia-smi --query-gpu=name,memory.total,power.limit --format=csv
