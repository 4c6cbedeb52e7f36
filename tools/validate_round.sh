# Round validation on one B200: GPU tests, smoke, bench lines of configs 2 / 3 / 5 and Path B, the reference arm, the ncu launch list of the step.
tag=${1:-r02aj}
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc $?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc $?"
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc $?"
timeout 600 python bench.py --config 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; echo "bench c3 rc $?"
timeout 900 python bench.py --config 5 --no-cpu-baseline > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err; echo "bench c5 rc $?"
timeout 900 python bench.py --path b --no-cpu-baseline > gpurun_out/${tag}_bench_pathb.json 2> gpurun_out/${tag}_bench_pathb.err; echo "bench path b rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc $?"
tail -3 gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_smoke.log
for f in n1 c3 c5 pathb reference; do tail -c 600 gpurun_out/${tag}_bench_$f.json; echo; done
