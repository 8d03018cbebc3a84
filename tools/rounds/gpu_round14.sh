#!/bin/bash
# parity margins of the tcgen05 path vs SYRK chain length; chunked M x M stage; configs[4] shapes on one GPU
TAG=${1:-r01t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x -k "chunked" > $OUT/pytest_chunked.log 2>&1; tail -5 $OUT/pytest_chunked.log
timeout 900 python tests/probes/parity_probe.py 2304,256,4 2304,256,4,1024 2304,256,4,512 2304,200,3 16384,256,4 16384,256,4,1024 16384,256,4,512 \
   32768,512,4 32768,512,4,1024 32768,1024,2 32768,1024,2,1024 32768,1024,2,512 > $OUT/parity_probe.jsonl 2> $OUT/parity_probe.err
cat $OUT/parity_probe.jsonl; tail -3 $OUT/parity_probe.err
timeout 300 python tools/tc_probe.py 1000000 1024 64 syrk > $OUT/syrk_chunk_sweep.jsonl 2>&1; cat $OUT/syrk_chunk_sweep.jsonl
timeout 900 python bench.py --n 131072 --m 4096 --l 128 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_big_1gpu.json 2> $OUT/bench_big_1gpu.err
head -c 3000 $OUT/bench_big_1gpu.json; tail -5 $OUT/bench_big_1gpu.err
nvidia-smi --query-gpu=memory.used --format=csv
