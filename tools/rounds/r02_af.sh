# MMA issue under elect.sync instead of `lane == 0` (no per-MMA waterfall loop for the descriptors): integer engine tests + bench
set -x
mkdir -p gpurun_out/r02af
timeout 300 python -m pytest tests/test_gpu_i8_engine.py -q -x > gpurun_out/r02af/pytest_i8.log 2>&1; tail -3 gpurun_out/r02af/pytest_i8.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lean > gpurun_out/r02af/bench.json 2> gpurun_out/r02af/bench.err; tail -c 300 gpurun_out/r02af/bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02af/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k: v for k, v in d['kernels_ms'].items() if v > 5}, d['clocks'])
PY
