# hand-written K3 adjoint + wide MMAs everywhere: full GPU suite, bench, M = 4096 ablation, full-size ncu captures
set -x
mkdir -p gpurun_out/r02g
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02g/pytest_gpu.log 2>&1; tail -3 gpurun_out/r02g/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02g/bench.json 2> gpurun_out/r02g/bench.err; tail -c 1200 gpurun_out/r02g/bench.json
timeout 900 python tests/probes/ablate_i8_probe.py 16384 4096 2 ,quad,kbwd,tn+nn+f32,syrk,scaledS,quad+kbwd+tn+nn+f32 > gpurun_out/r02g/ablate_m4096.jsonl 2> gpurun_out/r02g/ablate.err; cat gpurun_out/r02g/ablate_m4096.jsonl; tail -3 gpurun_out/r02g/ablate.err
# full captures at the bench size (one launch each, ~40 replays of the kernel)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scaled_i8_kernel --launch-skip 1 -c 1 -o gpurun_out/r02g/ncu_scaled_i8_wide_full python bench.py --steps 1 --warmup 0 --lean --no-cpu-baseline > gpurun_out/r02g/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:syrk_ -c 3 -o gpurun_out/r02g/ncu_syrk_full python bench.py --steps 1 --warmup 0 --lean --no-cpu-baseline > gpurun_out/r02g/ncu2.log 2>&1
ls -la gpurun_out/r02g
