#!/bin/bash
# isolated timing of the default attention kernel at three lengths + the attention / tower tests
tag=${1:-r02ag}
{
for T in 1025 1024 577 1025 1024; do echo -n "T=$T "; T=$T timeout 120 python tools/time_attn.py 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_parity.py -k "test_attention or clip_tower" -x -q 2>&1 | tail -3
} > gpurun_out/${tag}_attn_iso.log 2>&1
cat gpurun_out/${tag}_attn_iso.log
