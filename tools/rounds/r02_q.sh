# final validation: full GPU suite (incl. the M = 4096 parity tests), smoke, bench, full-size ncu of the scaled GEMM with the final epilogue
set -x
mkdir -p gpurun_out/r02q
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02q/pytest_gpu.log 2>&1; tail -4 gpurun_out/r02q/pytest_gpu.log
SVGP_FORCE_BUILD=0 timeout 300 python __graft_entry__.py smoke > gpurun_out/r02q/smoke.log 2>&1; tail -2 gpurun_out/r02q/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02q/bench.json 2> gpurun_out/r02q/bench.err; tail -c 600 gpurun_out/r02q/bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scaled_i8_kernel --launch-skip 1 -c 1 -o gpurun_out/r02q/ncu_scaled_i8_final python bench.py --steps 1 --warmup 0 --lean --no-cpu-baseline > gpurun_out/r02q/ncu1.log 2>&1
ls -la gpurun_out/r02q
