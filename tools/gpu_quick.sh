# One B200: the kernel parity tests and the K1 A/B at the engine's default policy (a short check after a K1 change).
# usage: bash tools/gpu_quick.sh [k1_ab specs...]
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_packed_gpu.py tests/test_multi_gpu.py -q -m gpu -x --timeout 600 2>&1 | tail -3
timeout 900 python tools/k1_ab.py ${@:-mandelmesh2048:d torus2048:d martin_cube1024:d p_key1024:d p_key1024_b2:d} 2>&1 | tee gpurun_out/k1_ab_quick.jsonl
