#!/bin/bash
# final profile of the round: ncu launch list of the bench command, full-scale DRAM traffic of the three tcgen05 kernels,
# --set full captures of the SCALED kernel (K-tile cache + 48-MMA chains) and of the vectorised K1 builder
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 1 --warmup 1 --n 262144 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max"
timeout 600 ncu --clock-control none --kernel-name-base mangled -k regex:tc_kernelILi2E --launch-skip 4 --launch-count 1 --csv --log-file $OUT/scaled_full_metrics.csv \
  --metrics $M python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_scaled_full.log 2>&1
grep -v "^==" $OUT/scaled_full_metrics.csv | cut -d, -f13- | tail -7
timeout 600 ncu --clock-control none --kernel-name-base mangled -k regex:tc_kernelILi0E --launch-skip 2 --launch-count 1 --csv --log-file $OUT/syrk_full_metrics.csv \
  --metrics $M python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_syrk_full.log 2>&1
grep -v "^==" $OUT/syrk_full_metrics.csv | cut -d, -f13- | tail -7
for spec in scaled_dots:29; do
  name=${spec%%:*}; skip=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_kernel --launch-skip $skip --launch-count 1 \
    -o $OUT/prof_$name python tools/tc_probe.py 131072 1024 16 > $OUT/full_$name.log 2>&1
  ncu -i $OUT/prof_$name.ncu-rep --page raw --csv > $OUT/prof_$name.raw.csv 2>/dev/null
  ncu -i $OUT/prof_$name.ncu-rep --page details --csv > $OUT/prof_$name.details.csv 2>/dev/null
  [ $(stat -c %s $OUT/prof_$name.ncu-rep) -gt 12000000 ] && rm -f $OUT/prof_$name.ncu-rep
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kernel_fwd_planes --launch-skip 1 --launch-count 1 \
  -o $OUT/prof_k1 python tools/tc_probe.py 1000000 1024 2 > $OUT/full_k1.log 2>&1
ncu -i $OUT/prof_k1.ncu-rep --page raw --csv > $OUT/prof_k1.raw.csv 2>/dev/null
ncu -i $OUT/prof_k1.ncu-rep --page details --csv > $OUT/prof_k1.details.csv 2>/dev/null
[ $(stat -c %s $OUT/prof_k1.ncu-rep) -gt 12000000 ] && rm -f $OUT/prof_k1.ncu-rep
ls -la $OUT | head -30
