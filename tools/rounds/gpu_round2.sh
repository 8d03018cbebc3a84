#!/bin/bash
# engine bring-up round: tcgen05 tests per k-block depth, full GPU suite, probe per depth, bench
TAG=${1:-r01c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for bk in 64 32; do
  SVGP_TC_BK=$bk timeout 600 python -m pytest tests/test_gpu_tc_engine.py -q -m gpu > $OUT/pytest_tc_bk$bk.log 2>&1; echo "rc=$?" >> $OUT/pytest_tc_bk$bk.log
done
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
for bk in 64 32; do
  SVGP_TC_BK=$bk timeout 600 python tools/tc_probe.py 262144 1024 16 > $OUT/tc_probe_bk$bk.jsonl 2> $OUT/tc_probe_bk$bk.err
done
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
for f in $OUT/pytest_tc_bk64.log $OUT/pytest_tc_bk32.log $OUT/pytest_gpu.log; do echo "== $f"; tail -15 $f; done
cat $OUT/tc_probe_bk*.jsonl
cat $OUT/bench.json | head -c 3500; tail -5 $OUT/bench.err
