#!/bin/bash
# after the accuracy fixes (consistent G_Kinv, S - Kinv stacking, 512-row SYRK chains, k-segmented SCALED chains)
TAG=${1:-r01v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python tests/probes/parity_probe.py 2304,256,4 16384,256,4 32768,512,4 32768,1024,2 65536,1024,2 32768,2048,1 > $OUT/parity_probe.jsonl 2> $OUT/parity_probe.err
cat $OUT/parity_probe.jsonl; tail -3 $OUT/parity_probe.err
SVGP_SCALED_KSEG=0 timeout 600 python tests/probes/parity_probe.py 32768,1024,2 > $OUT/parity_probe_kseg0.jsonl 2>/dev/null; cat $OUT/parity_probe_kseg0.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench.json | head -c 3500; tail -3 $OUT/bench.err
SVGP_SCALED_KSEG=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_kseg0.json 2> $OUT/bench_kseg0.err
python -c "
import json
for f in ('bench','bench_kseg0'):
    d=json.load(open('$OUT/%s.json'%f)); print(f, round(d['value']), d['ms_per_step'], {k:v for k,v in list(d['kernels_ms'].items())[:4]})
"
