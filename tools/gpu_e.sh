set -x
timeout 1800 python -m pytest tests/test_parity_gpu.py tests/test_multi_gpu.py tests/test_aux_gpu.py -q -m gpu -x --timeout 900 > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_e.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_e.json; tail -5 gpurun_out/bench_e.err
