# last validation of round 2: full GPU suite, smoke, bench, thirteen-pair SYRK cost at configs[4] shapes, ncu launch list,
# --set full captures of pass D (scaled_i8_kernel) and of the adjoint (three-digit, five-stage) pair SYRK
set -x
T=gpurun_out/r02final
mkdir -p $T
timeout 900 python -m pytest tests -m gpu -q > $T/pytest_gpu.log 2>&1; tail -4 $T/pytest_gpu.log
SVGP_FORCE_BUILD=0 timeout 300 python __graft_entry__.py smoke > $T/smoke.log 2>&1; tail -2 $T/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > $T/bench.json 2> $T/bench.err; tail -c 400 $T/bench.err; head -c 300 $T/bench.json; echo
timeout 300 python tests/probes/syrk_o4_timing.py 250000 4096 128 > $T/syrk_o4_timing.jsonl 2> $T/syrk_o4_timing.err; cat $T/syrk_o4_timing.jsonl; tail -2 $T/syrk_o4_timing.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $T/launches.csv python bench.py --steps 1 --warmup 1 --lean --no-cpu-baseline > $T/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scaled_i8_kernel --launch-skip 1 -c 1 -o $T/ncu_scaled_i8 python bench.py --steps 1 --warmup 0 --lean --no-cpu-baseline > $T/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:syrk_i8_pair_kernel --launch-skip 1 -c 1 -o $T/ncu_syrk_pair_d3 python bench.py --steps 1 --warmup 0 --lean --no-cpu-baseline > $T/ncu2.log 2>&1
ls -la $T
