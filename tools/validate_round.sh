set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc $?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.log 2>&1; echo "smoke rc $?"
timeout 600 python bench.py > gpurun_out/r02v_bench_n1.json 2> gpurun_out/r02v_bench_n1.err; echo "bench rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msqp_attention_tc -s 8 -c 1 -o gpurun_out/r02_msqp_attention_tc_full -f python tools/time_msqp.py > gpurun_out/ncu_msqp.log 2>&1; echo "ncu rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_d64_persist -s 2 -c 1 -o gpurun_out/r02v_attention_full -f python tools/run_attn.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu rc $?"
tail -3 gpurun_out/r02v_pytest.log; cat gpurun_out/r02v_smoke.log | tail -3
