# six-pair products (t + u <= 2) for the S - Kinv half of pass D: engine tests, parity at M = 256 .. 2048, bench
set -x
mkdir -p gpurun_out/r02ad
timeout 600 python -m pytest tests/test_gpu_i8_engine.py -q -x -k "scaled or bias or gemm_nn" > gpurun_out/r02ad/pytest_i8.log 2>&1; tail -12 gpurun_out/r02ad/pytest_i8.log
timeout 400 python tests/probes/parity_probe.py 16384,256,4 32768,512,4 32768,1024,2 32768,2048,2 65536,1024,8 > gpurun_out/r02ad/parity_d2.jsonl 2> gpurun_out/r02ad/parity_d2.err; cat gpurun_out/r02ad/parity_d2.jsonl
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lean > gpurun_out/r02ad/bench.json 2> gpurun_out/r02ad/bench.err; tail -c 300 gpurun_out/r02ad/bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02ad/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k: v for k, v in d['kernels_ms'].items() if v > 5}, d['clocks'])
PY
