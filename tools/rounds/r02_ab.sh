# pair SYRK with three leading digits: 40 KB stages (the raw operand's fourth plane is not loaded), five of them -- is the stage cycle
# latency what bounds the kernel?  engine tests + bench (svgp_syrk total of the two calls: 647 ms before)
set -x
mkdir -p gpurun_out/r02ab
timeout 600 python -m pytest tests/test_gpu_i8_engine.py -q -x -k "syrk or unbiased or pair_bias or three_leading" > gpurun_out/r02ab/pytest_i8.log 2>&1; tail -5 gpurun_out/r02ab/pytest_i8.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lean > gpurun_out/r02ab/bench.json 2> gpurun_out/r02ab/bench.err; tail -c 300 gpurun_out/r02ab/bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02ab/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k: v for k, v in d['kernels_ms'].items() if v > 5}, d['clocks'])
PY
