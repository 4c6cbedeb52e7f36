# A/B of the TMA fp32 epilogue (WG_GEMM_F32_TMA) of the pair GEMM inside the full step, same box
for v in 0 1 0 1; do
WG_GEMM_F32_TMA=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gather 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); k=d['kernels']['kernel_ms_per_step']
print('F32_TMA=$v', round(d['ms_per_step'],2), 'ms/step', round(d['value'],1), 'img/s | attn', k['attention_d64'], 'gemm2_bf16', k['gemm2_bf16'], 'f32', k.get('gemm2_f32'), 'ln', k['layernorm'], 'roofline', round(d['roofline']['frac'],3), d['clocks'].get('sm_mhz'))"
done
