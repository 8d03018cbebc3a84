# full bench line (device-timed value, end-to-end figure, roofline) of the final build; the CPU baseline is in r02_bench_final.json
mkdir -p gpurun_out/r02ah
timeout 70 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02ah/bench.json 2> gpurun_out/r02ah/bench.err; tail -c 200 gpurun_out/r02ah/bench.err; head -c 250 gpurun_out/r02ah/bench.json
