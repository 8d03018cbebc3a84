#!/bin/bash
# SYRK pipeline experiments: split barriers + packed-half transform, k-block depth, source-level stall profile, fp64 peaks
TAG=${1:-r01i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_tc_engine.py tests/test_gpu_primitives.py -x -q -m gpu > $OUT/pytest_tc.log 2>&1; echo "rc=$?" >> $OUT/pytest_tc.log
tail -5 $OUT/pytest_tc.log
(cd tools/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak) > $OUT/fp64_peak.jsonl 2>&1
cat $OUT/fp64_peak.jsonl
if grep -q "rc=0" $OUT/pytest_tc.log; then
  timeout 600 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/probe_h2.jsonl 2> $OUT/probe_h2.err
  SVGP_TC_DEBUG=2 timeout 600 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/probe_f32xf.jsonl 2> $OUT/probe_f32xf.err
  SVGP_TC_DEBUG=1 timeout 600 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/probe_noxf.jsonl 2> $OUT/probe_noxf.err
  SVGP_TC_BK=32 timeout 600 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/probe_bk32.jsonl 2> $OUT/probe_bk32.err
  SVGP_SYRK_SC=24576 timeout 600 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/probe_sc24k.jsonl 2> $OUT/probe_sc24k.err
  for f in h2 f32xf noxf bk32 sc24k; do echo $f; cat $OUT/probe_$f.jsonl | cut -c1-400; done
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_kernel --launch-skip 4 --launch-count 1 \
     -o $OUT/prof_syrk3 python tools/tc_probe.py 196608 1024 64 syrk > $OUT/full_syrk3.log 2>&1
  ncu -i $OUT/prof_syrk3.ncu-rep --page raw --csv > $OUT/prof_syrk3.raw.csv 2>/dev/null
  ncu -i $OUT/prof_syrk3.ncu-rep --page source --csv > $OUT/prof_syrk3.source.csv 2>/dev/null
  [ $(stat -c %s $OUT/prof_syrk3.ncu-rep) -gt 12000000 ] && rm -f $OUT/prof_syrk3.ncu-rep
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
  timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
  cat $OUT/bench.json | head -c 3500; tail -5 $OUT/bench.err
fi
