#!/bin/bash
# SYRK v2 bring-up: tcgen05 tests, full GPU suite, probe (two flush depths), bench
TAG=${1:-r01e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_tc_engine.py -x -q -m gpu > $OUT/pytest_tc.log 2>&1; echo "rc=$?" >> $OUT/pytest_tc.log
tail -5 $OUT/pytest_tc.log
if grep -q "rc=0" $OUT/pytest_tc.log; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
  timeout 600 python tools/tc_probe.py 262144 1024 16 > $OUT/tc_probe.jsonl 2> $OUT/tc_probe.err
  SVGP_SYRK_FLUSH=4 timeout 600 python tools/tc_probe.py 262144 1024 16 syrk > $OUT/tc_probe_flush4.jsonl 2> $OUT/tc_probe_flush4.err
  SVGP_SYRK_FLUSH=64 timeout 600 python tools/tc_probe.py 262144 1024 16 syrk > $OUT/tc_probe_flush64.jsonl 2> $OUT/tc_probe_flush64.err
  timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
  tail -5 $OUT/pytest_gpu.log
  cat $OUT/tc_probe*.jsonl | grep syrk
  cat $OUT/bench.json | head -c 3500; tail -5 $OUT/bench.err
fi
