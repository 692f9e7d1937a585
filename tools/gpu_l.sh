set -x
timeout 1800 python -m pytest tests/test_packed_gpu.py tests/test_parity_gpu.py tests/test_frontend_gpu.py tests/test_tolerance.py -q -m gpu -x --timeout 900 > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_l.log | cut -c1-300
timeout 900 python tools/k1_ab.py mandelmesh2048:d mandelmesh2048:d:S2M_K1_OPTIMISTIC=0 torus2048:d > gpurun_out/k1_opt.jsonl 2> gpurun_out/k1_opt.err; cat gpurun_out/k1_opt.jsonl | cut -c1-400; tail -3 gpurun_out/k1_opt.err
