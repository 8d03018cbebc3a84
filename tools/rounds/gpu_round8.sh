#!/bin/bash
# 2-GPU round: quick GPU tests, then the N=2 bench exactly as the driver launches it
TAG=${1:-r01k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_tc_engine.py tests/test_gpu_e2e.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench2.json 2> $OUT/bench2.err
echo "bench2 rc=$?"
cat $OUT/bench2.json | head -c 3000; tail -5 $OUT/bench2.err
