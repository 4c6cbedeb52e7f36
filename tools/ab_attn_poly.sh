# A/B of the polynomial-exponential share of the CLIP attention kernel inside the full (power-capped) step, same box
for v in 3 0 2 4 3 0 2 4; do
WG_ATTN_POLY=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gather 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); k=d['kernels']['kernel_ms_per_step']
print('POLY=$v', round(d['ms_per_step'],2), 'ms/step', round(d['value'],1), 'img/s | attn', k['attention_d64'], 'gemm2_bf16', k['gemm2_bf16'], 'f32', k.get('gemm2_f32'), d['clocks'].get('sm_mhz'), d['clocks'].get('power_w_max'))"
done
