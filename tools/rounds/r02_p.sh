# configs[4] shapes (M = 4096, L = 128) on 8 GPUs with the full SYRK (default for M > 2048) at a quarter of the rows per GPU
set -x
export PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True
mkdir -p gpurun_out/r02p
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --rows 250000 --inducing 4096 --channels 128 --steps 1 --warmup 1 --lean --no-cpu-baseline > gpurun_out/r02p/config5_quarterN_full_syrk.json 2> gpurun_out/r02p/config5.err
tail -3 gpurun_out/r02p/config5.err; head -c 2500 gpurun_out/r02p/config5_quarterN_full_syrk.json
