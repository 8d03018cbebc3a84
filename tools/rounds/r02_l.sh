# scaled GEMM epilogue: 8-column double-buffered chunks, early TMEM release, four arithmetic variants
set -x
mkdir -p gpurun_out/r02l
timeout 900 python -m pytest tests/test_gpu_i8_engine.py -x -q > gpurun_out/r02l/pytest_i8.log 2>&1; tail -5 gpurun_out/r02l/pytest_i8.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02l/bench.json 2> gpurun_out/r02l/bench.err; tail -c 1300 gpurun_out/r02l/bench.json
SVGP_I8_DEBUG=1 timeout 300 python bench.py --rows 262144 --steps 2 --warmup 1 --lean --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels_ms']; print('debug1 (epilogue skipped) scaled_i8', k['svgp_scaled_gemm_i8'])"
timeout 300 python bench.py --rows 262144 --steps 2 --warmup 1 --lean --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels_ms']; print('normal scaled_i8', k['svgp_scaled_gemm_i8'])"
