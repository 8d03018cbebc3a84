# vmax with 64 channels per pass, ball graph latency, actual gradient errors of the small configurations
set -x
mkdir -p gpurun_out/r02h
timeout 900 python -m pytest tests/test_gpu_i8_engine.py -x -q -k syrk > gpurun_out/r02h/pytest_i8.log 2>&1; tail -3 gpurun_out/r02h/pytest_i8.log
timeout 300 python tests/probes/small_grad_errors.py > gpurun_out/r02h/small_grad_errors.jsonl 2> gpurun_out/r02h/small_grad_errors.err; cat gpurun_out/r02h/small_grad_errors.jsonl
timeout 600 python tools/small_configs_timing.py > gpurun_out/r02h/small_configs.jsonl 2> gpurun_out/r02h/small_configs.err; cat gpurun_out/r02h/small_configs.jsonl
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02h/bench.json 2> gpurun_out/r02h/bench.err; tail -c 1500 gpurun_out/r02h/bench.json
