set -x
( nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 500 > gpurun_out/r02_probe_clocks.csv & echo $! > /tmp/smi.pid )
timeout 120 tools/micro/i8_mma_probe > gpurun_out/r02_i8_mma_probe.jsonl 2>&1; tail -2 gpurun_out/r02_i8_mma_probe.jsonl
kill $(cat /tmp/smi.pid)
# launch list of one full-size step (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_ncu_launch_list_fullsize.csv python bench.py --steps 1 --warmup 1 --lean --no-cpu-baseline > gpurun_out/r02_ncu_launch_bench.log 2>&1
# full captures of the two integer kernels at a reduced datapoint count (ncu replays every launch ~40 times)
ncu --set full --clock-control none --import-source on -k regex:scaled_i8_kernel -c 1 -o gpurun_out/r02_ncu_scaled_i8 python bench.py --rows 131072 --steps 1 --warmup 1 --lean --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:syrk_i8_pair_kernel -c 1 -o gpurun_out/r02_ncu_syrk_i8_pair python bench.py --rows 131072 --steps 1 --warmup 1 --lean --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
