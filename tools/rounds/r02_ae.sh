# thirteen-pair SYRK as its own instantiation (the ten-pair forward kernel had lost 12 % with the order-4 branches in it): SYRK engine
# tests + bench (svgp_syrk of the two calls: 608 ms before the order-4 code, 651 with it)
set -x
mkdir -p gpurun_out/r02ae
timeout 300 python -m pytest tests/test_gpu_i8_engine.py -q -x -k "syrk" > gpurun_out/r02ae/pytest_i8.log 2>&1; tail -3 gpurun_out/r02ae/pytest_i8.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lean > gpurun_out/r02ae/bench.json 2> gpurun_out/r02ae/bench.err; tail -c 300 gpurun_out/r02ae/bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02ae/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k: v for k, v in d['kernels_ms'].items() if v > 5}, d['clocks'])
PY
