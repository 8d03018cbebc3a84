#!/bin/bash
# One gpurun call: GPU parity tests, kernel probe, 1-GPU bench, ncu launch list and ncu --set full captures.
# Usage (on the box): bash tools/gpu_round.sh <tag> [stages]   stages = subset of "test probe bench launches full"
TAG=${1:-r01}
STAGES=${2:-"test probe bench launches full"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
for s in $STAGES; do
  case $s in
    test)
      timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log ;;
    probe)
      timeout 600 python tools/tc_probe.py 262144 1024 16 > $OUT/tc_probe.jsonl 2> $OUT/tc_probe.err ;;
    bench)
      timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ;;
    launches)
      # cold-cache, serialised per-launch times of one reduced-N step (shares, not absolutes)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
        python bench.py --steps 1 --warmup 1 --n 262144 --no-cpu-baseline > $OUT/launches_bench.log 2>&1 ;;
    full)
      # one launch of each tcgen05 mode (tc_probe launch order: 16 syrk, 4 rowquad full, 4 rowquad tri, 4 scaled, 4 scaled+dots)
      for spec in syrk:5 rowquad_full:17 rowquad_tri:21 scaled:25 scaled_dots:29; do
        name=${spec%%:*}; skip=${spec##*:}
        timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_kernel --launch-skip $skip --launch-count 1 \
          -o $OUT/prof_$name python tools/tc_probe.py 131072 1024 16 > $OUT/full_$name.log 2>&1
        ncu -i $OUT/prof_$name.ncu-rep --page raw --csv > $OUT/prof_$name.raw.csv 2>/dev/null
        ncu -i $OUT/prof_$name.ncu-rep --page details --csv > $OUT/prof_$name.details.csv 2>/dev/null
        ncu -i $OUT/prof_$name.ncu-rep --page source --csv > $OUT/prof_$name.source.csv 2>/dev/null
        [ $(stat -c %s $OUT/prof_$name.ncu-rep) -gt 12000000 ] && rm -f $OUT/prof_$name.ncu-rep
      done
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:kernel_fwd_kernel --launch-skip 1 --launch-count 1 \
        -o $OUT/prof_k1 python tools/tc_probe.py 262144 1024 2 > $OUT/full_k1.log 2>&1
      ncu -i $OUT/prof_k1.ncu-rep --page raw --csv > $OUT/prof_k1.raw.csv 2>/dev/null
      ncu -i $OUT/prof_k1.ncu-rep --page details --csv > $OUT/prof_k1.details.csv 2>/dev/null
      [ $(stat -c %s $OUT/prof_k1.ncu-rep) -gt 12000000 ] && rm -f $OUT/prof_k1.ncu-rep ;;
  esac
done
ls -la $OUT
du -sm gpurun_out
tail -3 $OUT/pytest_gpu.log 2>/dev/null
cat $OUT/bench.json 2>/dev/null | head -c 3000
