# whole -m gpu suite (all failures), then the benches with the new record
set -x
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_b.log
timeout 600 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_b.json; tail -5 gpurun_out/bench_b.err
S2M_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --no-other-workloads --no-verify --steps 1 --warmup 3 > /dev/null 2> gpurun_out/trace_b.txt; tail -120 gpurun_out/trace_b.txt | head -150 > gpurun_out/trace_b_tail.txt
