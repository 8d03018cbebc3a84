# float64-evaluated small K_nm: full GPU suite, smoke, actual small-configuration gradient errors
set -x
mkdir -p gpurun_out/r02i
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02i/pytest_gpu.log 2>&1; tail -3 gpurun_out/r02i/pytest_gpu.log
SVGP_FORCE_BUILD=0 timeout 300 python __graft_entry__.py smoke > gpurun_out/r02i/smoke.log 2>&1; tail -3 gpurun_out/r02i/smoke.log
timeout 300 python tests/probes/small_grad_errors.py > gpurun_out/r02i/small_grad_errors.jsonl 2> gpurun_out/r02i/small_grad_errors.err; cat gpurun_out/r02i/small_grad_errors.jsonl
timeout 600 python tools/small_configs_timing.py > gpurun_out/r02i/small_configs.jsonl 2> gpurun_out/r02i/small_configs.err; cat gpurun_out/r02i/small_configs.jsonl
