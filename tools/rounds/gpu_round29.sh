#!/bin/bash
TAG=${1:-r02n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_e2e.py -m gpu -q -x --timeout 150 -k "kernel_fwd_bwd or sweep" > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
python -c "
import json
d=json.load(open('$OUT/bench.json')); print(round(d['value']), round(d['ms_per_step'],1), {k:v for k,v in d['kernels_ms'].items() if 'kernel' in k or 'gemm_tn' in k or 'gemm_f32' in k})
" || tail -3 $OUT/bench.err
timeout 200 python tests/probes/parity_probe.py 32768,1024,2 16384,256,4 2>/dev/null | tee $OUT/parity.jsonl
