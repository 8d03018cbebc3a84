#!/bin/bash
TAG=${1:-r01y}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
SVGP_SYRK_BIAS=0 timeout 600 python tests/probes/accum_probe.py 32768 1024 2 2>/dev/null | grep syrk > $OUT/accum_syrk_raw.jsonl; cat $OUT/accum_syrk_raw.jsonl
timeout 600 python tests/probes/accum_probe.py 32768 1024 2 2>/dev/null | grep syrk > $OUT/accum_syrk_corrected.jsonl; cat $OUT/accum_syrk_corrected.jsonl
timeout 900 python tests/probes/ablate_probe.py 32768 1024 2 > $OUT/ablate.jsonl 2> $OUT/ablate.err; cat $OUT/ablate.jsonl; tail -3 $OUT/ablate.err
SVGP_SCALED_KSEG=4 timeout 900 python tests/probes/ablate_probe.py 32768 1024 2 > $OUT/ablate_kseg4.jsonl 2> $OUT/ablate.err; cat $OUT/ablate_kseg4.jsonl
timeout 900 python tests/probes/parity_probe.py 2304,256,4 16384,256,4 32768,512,4 32768,1024,2 65536,1024,2 32768,2048,2 > $OUT/parity_probe.jsonl 2> $OUT/parity_probe.err; cat $OUT/parity_probe.jsonl
timeout 600 python tests/probes/parity_fullsize.py 524288 1024 2 > $OUT/parity_fullsize.jsonl 2> $OUT/parity_fullsize.err; cat $OUT/parity_fullsize.jsonl; tail -2 $OUT/parity_fullsize.err
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
python -c "
import json
d=json.load(open('$OUT/bench.json')); print(round(d['value']), round(d['ms_per_step'],1), d['kernels_ms'])
"
