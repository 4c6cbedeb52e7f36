for v in 1 0 1 0; do
WG_ATTN_EXTRA_KEY=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gather 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); k=d['kernels']['kernel_ms_per_step']
print('EXTRA=$v', round(d['ms_per_step'],2), 'attn', k['attention_d64'], 'gemm2_bf16', k['gemm2_bf16'], 'f32', k['gemm2_f32'], d['clocks']['sm_mhz'])"
done
