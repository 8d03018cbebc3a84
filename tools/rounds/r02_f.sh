# wide (N = 256) MMAs in the scaled GEMM, SYRK diagonal-block split: correctness of every variant, A/B timings, parity incl. M = 4096
set -x
mkdir -p gpurun_out/r02f
timeout 900 python -m pytest tests/test_gpu_i8_engine.py -x -q > gpurun_out/r02f/pytest_i8.log 2>&1; tail -3 gpurun_out/r02f/pytest_i8.log
for wide in 1 0; do for split in 1 0; do
SVGP_I8_WIDE=$wide SVGP_I8_SYRK_SPLIT=$split timeout 300 python bench.py --rows 262144 --steps 2 --warmup 1 --lean --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels_ms']; print(json.dumps(dict(wide=$wide, split=$split, scaled_i8=k['svgp_scaled_gemm_i8'], syrk=k['svgp_syrk'], step=d['ms_per_step'])))" | tee -a gpurun_out/r02f/ab.jsonl
done; done
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02f/bench.json 2> gpurun_out/r02f/bench.err; tail -c 1500 gpurun_out/r02f/bench.json
timeout 900 python tests/probes/parity_probe.py 32768,1024,2 32768,2048,2 16384,4096,2 > gpurun_out/r02f/parity.jsonl 2> gpurun_out/r02f/parity.err; cat gpurun_out/r02f/parity.jsonl; tail -3 gpurun_out/r02f/parity.err
