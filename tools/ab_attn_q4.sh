#!/bin/bash
# A/B of the four-CTAs-per-SM attention kernel (WG_ATTN_Q4=1) against the persistent two-CTA kernel, isolated, same box.
mkdir -p gpurun_out
{
for T in 1025 1024 577; do
for q4 in 0 1; do for poly in 0 2 3; do
  echo -n "T=$T Q4=$q4 "; T=$T WG_ATTN_Q4=$q4 WG_ATTN_POLY=$poly timeout 120 python tools/time_attn.py 2>&1 | tail -1
done; done; done
WG_ATTN_Q4=1 timeout 600 python -m pytest tests/test_gpu_parity.py -k "test_attention" -x -q 2>&1 | tail -5
} > gpurun_out/${1:-r02aa}_attn_q4.log 2>&1
tail -30 gpurun_out/${1:-r02aa}_attn_q4.log
