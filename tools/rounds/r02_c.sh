set -x
python -m pytest tests/test_gpu_i8_engine.py tests/test_gpu_e2e.py -q -x 2>&1 | tail -6
python tests/probes/parity_probe.py 32768,1024,2 32768,2048,2 2>/dev/null
SVGP_I8_PAIR=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('single', d['value'], d['ms_per_step']); print(d['kernels_ms'])"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null > gpurun_out/r02_bench_pair_v1.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_pair_v1.json')); print('pair', d['value'], d['ms_per_step']); print(d['kernels_ms'])"
