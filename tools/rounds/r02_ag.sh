# elect.sync guards for the fp16 engine's MMA issue and for all TMA producers as well: full GPU suite + bench
set -x
mkdir -p gpurun_out/r02ag
timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/r02ag/pytest_gpu.log 2>&1; tail -3 gpurun_out/r02ag/pytest_gpu.log
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lean > gpurun_out/r02ag/bench.json 2> gpurun_out/r02ag/bench.err; tail -c 300 gpurun_out/r02ag/bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02ag/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k: v for k, v in d['kernels_ms'].items() if v > 5}, d['clocks'])
PY
