# expected value of the dropped digit-plane pairs inside the integer scaled GEMM (svgp_i8_pair_bias): engine tests, parity at
# M = 4096 / 2048 / 1024, bench
set -x
mkdir -p gpurun_out/r02aa
timeout 600 python -m pytest tests/test_gpu_i8_engine.py -q -x > gpurun_out/r02aa/pytest_i8.log 2>&1; tail -15 gpurun_out/r02aa/pytest_i8.log
timeout 400 python tests/probes/parity_probe.py 16384,4096,2 32768,2048,2 32768,1024,2 > gpurun_out/r02aa/parity_debias.jsonl 2> gpurun_out/r02aa/parity_debias.err; cat gpurun_out/r02aa/parity_debias.jsonl
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lean > gpurun_out/r02aa/bench_debias.json 2> gpurun_out/r02aa/bench_debias.err; tail -c 300 gpurun_out/r02aa/bench_debias.err; head -c 300 gpurun_out/r02aa/bench_debias.json
