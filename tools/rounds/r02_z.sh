# three-leading-digit products (SVGP_IMPL_TC_I8_D3 / nfull) for the adjoint SYRK and the S - Kinv family of pass D:
# engine tests against the digit-exact emulation, parity at M = 1024 / 2048 / 4096 with and without, bench with and without
set -x
mkdir -p gpurun_out/r02z
timeout 600 python -m pytest tests/test_gpu_i8_engine.py -q -x > gpurun_out/r02z/pytest_i8.log 2>&1; tail -3 gpurun_out/r02z/pytest_i8.log
timeout 400 python tests/probes/parity_probe.py 32768,1024,2 32768,2048,2 16384,4096,2 65536,1024,8 > gpurun_out/r02z/parity_d3.jsonl 2> gpurun_out/r02z/parity_d3.err; cat gpurun_out/r02z/parity_d3.jsonl
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02z/bench_d3.json 2> gpurun_out/r02z/bench_d3.err; tail -c 300 gpurun_out/r02z/bench_d3.err; head -c 400 gpurun_out/r02z/bench_d3.json
