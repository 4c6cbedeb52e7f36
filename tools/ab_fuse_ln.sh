# A/B of the fused residual + LayerNorm epilogue in the full step (same box): WG_CLIP_FUSE_LN = 0 / 1 / 2
for v in 0 1 2 0 1 2; do
WG_CLIP_FUSE_LN=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gather 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); k=d['kernels']['kernel_ms_per_step']
print('FUSE_LN=$v', round(d['ms_per_step'],2), 'ms/step', round(d['value'],1), 'img/s | attn', k['attention_d64'], 'gemm2_bf16', k['gemm2_bf16'], 'f32', k.get('gemm2_f32'), 'f32_ln', k.get('gemm2_f32_ln'), 'ln', k['layernorm'], d['clocks']['sm_mhz'])"
done
