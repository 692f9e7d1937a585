set -x
timeout 900 python -m pytest tests/test_packed_gpu.py -q -m gpu -x --timeout 900 2>&1 | tail -5
timeout 900 python tools/k1_ab.py mandelmesh2048:d mandelmesh2048:d:S2M_K1_OPTIMISTIC=0 > gpurun_out/k1_opt2.jsonl 2> gpurun_out/k1_opt2.err; cat gpurun_out/k1_opt2.jsonl | cut -c1-400; tail -3 gpurun_out/k1_opt2.err
B="python bench.py --steps 1 --warmup 1 --no-verify --no-cpu-baseline --no-other-workloads"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2m_k1_slab -s 17 -c 1 -o gpurun_out/r02_k1_mandel_opt2 $B > /dev/null 2>&1
