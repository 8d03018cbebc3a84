# launch list of one full-size step with the final kernels (shares), full capture of the triangular row-quad kernel
set -x
mkdir -p gpurun_out/r02j
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02j/launches.csv python bench.py --steps 1 --warmup 1 --lean --no-cpu-baseline > gpurun_out/r02j/ncu_bench.log 2>&1
tail -c 300 gpurun_out/r02j/ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_kernel --launch-skip 1 -c 1 -o gpurun_out/r02j/ncu_quad_full python bench.py --steps 1 --warmup 0 --lean --no-cpu-baseline > gpurun_out/r02j/ncu2.log 2>&1
ls -la gpurun_out/r02j
# the M sweep of configs[3] at N = 1e6, L = 64 on one GPU with the final kernels
for m in 256 512 2048; do
  timeout 900 python bench.py --inducing $m --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02j/sweep_m$m.json 2> gpurun_out/r02j/sweep_m$m.err
  tail -c 300 gpurun_out/r02j/sweep_m$m.json
done
