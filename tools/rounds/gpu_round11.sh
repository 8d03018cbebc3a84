#!/bin/bash
# DRAM traffic of the dominant kernel at full scale (for roofline.traffic) + the M sweep of BASELINE configs[3]
TAG=${1:-r01q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
# tc_kernel<2,...> = SCALED: 1 launch per step; skip the warm-up step's launch
timeout 900 ncu --clock-control none --kernel-name-base mangled -k regex:tc_kernelILi2E --launch-skip 4 --launch-count 1 --csv --log-file $OUT/scaled_full_metrics.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_scaled_full.log 2>&1
grep -v "^==" $OUT/scaled_full_metrics.csv | cut -d, -f17- | tail -7
timeout 900 ncu --clock-control none --kernel-name-base mangled -k regex:tc_kernelILi1E --launch-skip 3 --launch-count 1 --csv --log-file $OUT/quad_full_metrics.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_quad_full.log 2>&1
grep -v "^==" $OUT/quad_full_metrics.csv | cut -d, -f17- | tail -7
for m in ; do
  timeout 900 python bench.py --m $m --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_m$m.json 2> $OUT/bench_m$m.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_m$m.json"))
    print("M=$m", round(d["value"]), "dp/s", round(d["ms_per_step"], 1), "ms", {k: round(v) for k, v in list(d["kernels_ms"].items())[:6]})
except Exception as e:
    print("M=$m failed", e); print(open("$OUT/bench_m$m.err").read()[-1500:])
PY
done
