{
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -k "test_attention and not 16-1025 and not 40-512" -x -q 2>&1 | tail -8; echo "memcheck rc ${PIPESTATUS[0]}"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); print(round(d['value'],1), d['kernels']['attention_d64'])"
} > gpurun_out/r02ak_sanitizer.log 2>&1
cat gpurun_out/r02ak_sanitizer.log
