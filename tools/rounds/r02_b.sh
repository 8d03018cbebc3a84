set -x
python -m pytest tests/test_gpu_i8_engine.py -q 2>&1 | tail -15
python tests/probes/parity_probe.py 16384,256,4 32768,512,4 32768,1024,2 32768,2048,2 > gpurun_out/r02_parity_i8_v6.jsonl 2>gpurun_out/r02_parity_i8_v6.err; tail -3 gpurun_out/r02_parity_i8_v6.err; cat gpurun_out/r02_parity_i8_v6.jsonl
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_i8_v6.json 2>gpurun_out/r02_bench_i8_v6.err; tail -3 gpurun_out/r02_bench_i8_v6.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_i8_v6.json')); print(d['value'], d['ms_per_step']); print(d['kernels_ms'])"
