# thirteen-pair forward SYRK (order-4 items on the pair kernel, M > 2048) + the five-stage three-digit pair SYRK: engine tests,
# parity at M = 4096 (probe inputs and the test's inputs), timing of the SYRK at M = 4096
set -x
mkdir -p gpurun_out/r02ac
timeout 600 python -m pytest tests/test_gpu_i8_engine.py -q -x -k "syrk" > gpurun_out/r02ac/pytest_i8.log 2>&1; tail -5 gpurun_out/r02ac/pytest_i8.log
timeout 400 python tests/probes/parity_probe.py 16384,4096,2 > gpurun_out/r02ac/parity_o4.jsonl 2> gpurun_out/r02ac/parity_o4.err; cat gpurun_out/r02ac/parity_o4.jsonl
timeout 400 python -m pytest tests/test_gpu_fullsize.py -q -k "configs4" > gpurun_out/r02ac/pytest_m4096.log 2>&1; tail -5 gpurun_out/r02ac/pytest_m4096.log; grep -o "{'N': 16384[^}]*}" gpurun_out/r02ac/pytest_m4096.log | tail -1
