# one 8-GPU box: strong scaling of the N = 1e6 sweep point at 1 / 2 / 4 / 8 GPUs with the final kernels (wide MMAs, diagonal split, hand-written K3 adjoint)
set -x
export PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True
mkdir -p gpurun_out/r02s2
python bench.py --strong --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02s2/strong_1gpu.json 2> gpurun_out/r02s2/strong_1gpu.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --strong --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02s2/strong_${n}gpu.json 2> gpurun_out/r02s2/strong_${n}gpu.err
  tail -c 400 gpurun_out/r02s2/strong_${n}gpu.json
done
# weak scaling at 8 GPUs (the driver's own scaling run uses the default = weak): 1e6 rows per GPU
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02s2/weak_8gpu.json 2> gpurun_out/r02s2/weak_8gpu.err
tail -c 400 gpurun_out/r02s2/weak_8gpu.json
