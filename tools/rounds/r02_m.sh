# full (both triangles, averaged) SYRK: exactness, parity at M = 4096 / 2048 with and without, cost at the headline size
set -x
mkdir -p gpurun_out/r02m
timeout 900 python -m pytest tests/test_gpu_i8_engine.py -x -q -k syrk > gpurun_out/r02m/pytest_i8.log 2>&1; tail -3 gpurun_out/r02m/pytest_i8.log
for full in 1 0; do
SVGP_I8_SYRK_FULL=$full timeout 900 python tests/probes/parity_probe.py 16384,4096,2 32768,2048,2 2> /dev/null | sed "s/^/{\"full\": $full, /; s/, {/, /" | tee -a gpurun_out/r02m/parity_full.jsonl
done
SVGP_I8_SYRK_FULL=1 timeout 300 python bench.py --rows 262144 --steps 2 --warmup 1 --lean --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels_ms']; print(json.dumps(dict(full=1, syrk=k['svgp_syrk'], step=d['ms_per_step'])))" | tee -a gpurun_out/r02m/cost.jsonl
