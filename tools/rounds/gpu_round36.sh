#!/bin/bash
# one shot: SYRK with split accumulators (cross terms in the second TMEM buffer)
TAG=${1:-r02y}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
export SVGP_SYRK_SPLIT_ACC=1
timeout 40 python -m pytest tests/test_gpu_tc_engine.py -m gpu -q -x --timeout 30 -k "syrk" 2>&1 | tail -2 | tee $OUT/pytest.log
timeout 45 python tests/probes/parity_probe.py 32768,1024,2 32768,2048,2 2>/dev/null | tee $OUT/parity_split.jsonl
timeout 30 python tools/tc_probe.py 1000000 1024 64 syrk 2>&1 | grep '"chunk": 512\|rror' | tee $OUT/syrk_split.jsonl
SVGP_SYRK_BIAS=0 timeout 50 python tests/probes/accum_probe.py 32768 1024 2 2>/dev/null | grep "syrk\[0\]" | head -4 | tee $OUT/accum_split_raw.jsonl
