set -x
mkdir -p gpurun_out/r02o
SVGP_I8_SYRK_FULL=1 timeout 900 python tests/probes/ablate_i8_probe.py 16384 4096 2 dotsA,outA > gpurun_out/r02o/ablate_m4096_dots_out.jsonl 2> gpurun_out/r02o/ablate.err; cat gpurun_out/r02o/ablate_m4096_dots_out.jsonl; tail -3 gpurun_out/r02o/ablate.err
