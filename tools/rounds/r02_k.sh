# final validation of the tree + the integer engine at M = 4096 against its digit-exact emulation
set -x
mkdir -p gpurun_out/r02k
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02k/pytest_gpu.log 2>&1; tail -4 gpurun_out/r02k/pytest_gpu.log
SVGP_FORCE_BUILD=0 timeout 300 python __graft_entry__.py smoke > gpurun_out/r02k/smoke.log 2>&1; tail -2 gpurun_out/r02k/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02k/bench.json 2> gpurun_out/r02k/bench.err; tail -c 600 gpurun_out/r02k/bench.json
