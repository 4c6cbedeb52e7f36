#!/bin/bash
# A/B of WG_ATTN_Q4 inside the whole step (same box), isolated timing of the new default, attention tests, ncu capture of the q4 kernel
tag=${1:-r02ab}
{
for T in 1025 1024; do echo -n "T=$T default "; T=$T timeout 120 python tools/time_attn.py 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_parity.py -k "test_attention or clip_tower or full_batch" -x -q 2>&1 | tail -3
for v in 1 0 1 0; do
WG_ATTN_Q4=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gather 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); k=d['kernels']['kernel_ms_per_step']
print('Q4=$v', round(d['ms_per_step'],2), round(d['value'],1), 'attn', k['attention_d64'], 'gemm2_bf16', k['gemm2_bf16'], 'f32', k['gemm2_f32'], d['clocks']['sm_mhz'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_d64_q4 -s 2 -c 1 -o gpurun_out/${tag}_attention_q4_full -f python tools/run_attn.py > gpurun_out/ncu_attn_q4.log 2>&1; echo "ncu rc $?"
} > gpurun_out/${tag}_q4_step.log 2>&1
cat gpurun_out/${tag}_q4_step.log
