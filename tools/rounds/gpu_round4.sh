#!/bin/bash
# full-scale SYRK investigation: timing with/without the operand transform, ncu counters of one launch, DGEMM test
TAG=${1:-r01f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_primitives.py -x -q -m gpu -k linalg > $OUT/pytest_linalg.log 2>&1; echo "rc=$?" >> $OUT/pytest_linalg.log
tail -3 $OUT/pytest_linalg.log
timeout 600 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/probe_full.jsonl 2> $OUT/probe_full.err
SVGP_TC_DEBUG=1 timeout 600 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/probe_full_noxf.jsonl 2> $OUT/probe_full_noxf.err
timeout 900 ncu --clock-control none -k regex:tc_kernel --launch-skip 4 --launch-count 1 --csv --log-file $OUT/syrk_full_metrics.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,lts__t_sectors_srcunit_tex.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum \
  python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/ncu_syrk_full.log 2>&1
timeout 300 python tools/tc_probe.py 262144 1024 16 > $OUT/tc_probe.jsonl 2> $OUT/tc_probe.err
cat $OUT/probe_full.jsonl $OUT/probe_full_noxf.jsonl
grep -v "^==" $OUT/syrk_full_metrics.csv | cut -d, -f5,13- | head -30
grep "bmm64\|chol\|trinv" $OUT/tc_probe.jsonl
