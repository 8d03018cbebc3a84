for pair in 0 1; do for dbg in 0 1; do
SVGP_I8_PAIR=$pair SVGP_I8_DEBUG=$dbg python bench.py --rows 262144 --steps 2 --warmup 1 --lean --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels_ms']; print('pair=$pair dbg=$dbg scaled_i8', k['svgp_scaled_gemm_i8'], 'syrk', k['svgp_syrk'])"
done; done
