set -x
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_packed_gpu.py tests/test_frontend_gpu.py -q -m gpu -x --timeout 900 2>&1 | tail -4
timeout 1500 python tools/k1_ab.py mandelmesh2048:d torus2048:d torus2048:d:S2M_K1_ZPT=32 torus2048:d:S2M_K1_ZPT=16,S2M_K1_ROWS=1 torus2048:d:S2M_K1_ZPT=32,S2M_K1_MINBLOCKS=5 \
  martin_cube512:d martin_cube512:d:S2M_K1_ZPT=4 martin_cube512:d:S2M_K1_MINBLOCKS=5 martin_cube512:d:S2M_K1_MINBLOCKS=3 \
  p_key1024_b2:d p_key1024_b2:d:S2M_K1_ZPT=4 p_key1024_b2:d:S2M_K1_ZPT=8 p_key1024_b2:d:S2M_K1_MINBLOCKS=5 p_key1024_b2:d:S2M_K1_MINBLOCKS=6 p_key1024:d \
  > gpurun_out/k1_ab_r02b.jsonl 2> gpurun_out/k1_ab_r02b.err; cat gpurun_out/k1_ab_r02b.jsonl | cut -c1-330; tail -3 gpurun_out/k1_ab_r02b.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_g.err
