# where does dhyp at M = 4096 come from?  (full SYRK on)
set -x
mkdir -p gpurun_out/r02n
SVGP_I8_SYRK_FULL=1 timeout 900 python tests/probes/ablate_i8_probe.py 16384 4096 2 ,scaledA,scaledA+kbwd,ktrue,scaledA+syrk+quad+kbwd+tn+nn+f32 > gpurun_out/r02n/ablate_m4096_full.jsonl 2> gpurun_out/r02n/ablate.err; cat gpurun_out/r02n/ablate_m4096_full.jsonl; tail -3 gpurun_out/r02n/ablate.err
