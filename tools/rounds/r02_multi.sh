# one 8-GPU box: strong scaling of the N = 1e6 sweep point at 1 / 2 / 4 / 8 GPUs, then configs[4] (lean: 1 warm-up + 2 timed steps)
set -x
export PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True
python bench.py --strong --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02_strong_1gpu.json 2> gpurun_out/r02_strong_1gpu.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --strong --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02_strong_${n}gpu.json 2> gpurun_out/r02_strong_${n}gpu.err
  tail -c 600 gpurun_out/r02_strong_${n}gpu.json
done
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --rows 1000000 --inducing 4096 --channels 128 --steps 2 --warmup 1 --lean --no-cpu-baseline > gpurun_out/r02_bench_config5_8gpu.json 2> gpurun_out/r02_bench_config5_8gpu.err
tail -5 gpurun_out/r02_bench_config5_8gpu.err
cat gpurun_out/r02_bench_config5_8gpu.json | head -c 3000
