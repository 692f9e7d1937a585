# K2 on the consumer stream + segment-count K3: parity suite, bench, trace
set -x
timeout 1800 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_d.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_d.json; tail -5 gpurun_out/bench_d.err
S2M_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --no-other-workloads --no-verify --steps 1 --warmup 3 > /dev/null 2> gpurun_out/trace_d.txt
