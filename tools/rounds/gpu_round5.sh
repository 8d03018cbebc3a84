#!/bin/bash
# super-chunked SYRK: tests, full-scale probe + ncu counters, bench
TAG=${1:-r01g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_tc_engine.py -x -q -m gpu > $OUT/pytest_tc.log 2>&1; echo "rc=$?" >> $OUT/pytest_tc.log
tail -5 $OUT/pytest_tc.log
if grep -q "rc=0" $OUT/pytest_tc.log; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
  timeout 600 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/probe_full.jsonl 2> $OUT/probe_full.err
  SVGP_SYRK_SC=6144 timeout 600 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/probe_full_sc6k.jsonl 2> $OUT/probe_full_sc6k.err
  SVGP_SYRK_SC=24576 timeout 600 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/probe_full_sc24k.jsonl 2> $OUT/probe_full_sc24k.err
  timeout 900 ncu --clock-control none -k regex:tc_kernel --launch-skip 4 --launch-count 1 --csv --log-file $OUT/syrk_full_metrics.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max \
    python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/ncu_syrk_full.log 2>&1
  timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
  cat $OUT/probe_full*.jsonl
  grep -v "^==" $OUT/syrk_full_metrics.csv | cut -d, -f13- | head -30
  cat $OUT/bench.json | head -c 3500; tail -5 $OUT/bench.err
fi
