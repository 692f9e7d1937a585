set -x
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_packed_gpu.py tests/test_multi_gpu.py -q -m gpu -x --timeout 900 2>&1 | tail -4
timeout 1500 python tools/k1_ab.py mandelmesh2048:d:S2M_K1_ZPT=1 mandelmesh2048:d:S2M_K1_ZPT=2 mandelmesh2048:d:S2M_K1_ZPT=4 mandelmesh2048:d:S2M_K1_ZPT=8 mandelmesh2048:d:S2M_K1_ZPT=16 \
  torus2048:d:S2M_K1_ZPT=1 torus2048:d:S2M_K1_ZPT=4 torus2048:d:S2M_K1_ZPT=8 torus2048:d:S2M_K1_ZPT=16 \
  martin_cube512:d:S2M_K1_ZPT=1 martin_cube512:d:S2M_K1_ZPT=4 martin_cube512:d:S2M_K1_ROWS=1 martin_cube512:d:S2M_K1_ROWS=1,S2M_K1_UNROLL=1 martin_cube512:d:S2M_K1_MINBLOCKS=2 martin_cube512:d:S2M_K1_MINBLOCKS=3 martin_cube512:d:S2M_K1_ROWS=1,S2M_K1_MINBLOCKS=3 martin_cube512:d:S2M_K1_ROWS=1,S2M_K1_MINBLOCKS=4 martin_cube512:d:S2M_K1_ROWS=1,S2M_K1_UNROLL=1,S2M_K1_MINBLOCKS=4 \
  p_key1024_b2:d:S2M_K1_ZPT=1 p_key1024_b2:d:S2M_K1_ZPT=4 p_key1024_b2:d:S2M_K1_ROWS=1 p_key1024_b2:d:S2M_K1_MINBLOCKS=3 p_key1024_b2:d:S2M_K1_MINBLOCKS=4 p_key1024_b2:d:S2M_K1_ROWS=1,S2M_K1_MINBLOCKS=4 p_key1024_b2:d:S2M_K1_ROWS=1,S2M_K1_UNROLL=1,S2M_K1_MINBLOCKS=5 \
  > gpurun_out/k1_ab_r02.jsonl 2> gpurun_out/k1_ab_r02.err; cat gpurun_out/k1_ab_r02.jsonl | cut -c1-330; tail -3 gpurun_out/k1_ab_r02.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_f.err
