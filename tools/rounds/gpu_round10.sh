#!/bin/bash
# cluster-multicast SYRK bring-up (tight timeouts: a protocol bug shows up as a hang)
TAG=${1:-r01n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 240 python -m pytest tests/test_gpu_tc_engine.py -x -q -m gpu -k syrk > $OUT/pytest_syrk.log 2>&1; echo "rc=$?" >> $OUT/pytest_syrk.log
tail -6 $OUT/pytest_syrk.log
if grep -q "rc=0" $OUT/pytest_syrk.log; then
  timeout 300 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/probe_mc.jsonl 2> $OUT/probe_mc.err
  SVGP_SYRK_CLUSTER=1 timeout 300 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/probe_nomc.jsonl 2> $OUT/probe_nomc.err
  for f in mc nomc; do echo $f; cut -c1-200 $OUT/probe_$f.jsonl; done
  timeout 600 ncu --clock-control none -k regex:tc_kernel --launch-skip 4 --launch-count 1 --csv --log-file $OUT/syrk_full_metrics.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max \
    python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/ncu_syrk_full.log 2>&1
  grep -v "^==" $OUT/syrk_full_metrics.csv | cut -d, -f13- | tail -8
fi
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench.json | head -c 3000; tail -3 $OUT/bench.err
