# A/B of the decoder's per-image layer-0 projections (WG_DEC_SHARE_L0) at 3 and 12 [SEG] per image, same box
for cfg in 2 3; do for v in 0 1 0 1; do
WG_DEC_SHARE_L0=$v timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-gather 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); k=d['kernels']['kernel_ms_per_step']
print('config $cfg SHARE_L0=$v', round(d['ms_per_step'],2), 'ms/step', round(d['value'],1), 'img/s | f32_bn128', k.get('gemm_f32_bn128'), 't2i', k['dec_t2i_attention'], 'i2t', k['dec_i2t_attention'], 'expand', k['dec_expand_keys'], 'tok', k['dec_token'], 'bf16ln', k['gemm_bf16ln_bn256'])"
done; done
