# one-GPU validation of the packed K1: parity (both variants), A/B timings, the default bench line, then the whole GPU suite
set -x
mkdir -p gpurun_out
python -m pytest tests/test_packed_gpu.py -x -q -m gpu > gpurun_out/packed_tests_v1.log 2>&1; echo "v1 rc=$?"; tail -3 gpurun_out/packed_tests_v1.log
S2M_TEST_PACKED_VARIANT=2 python -m pytest tests/test_packed_gpu.py -x -q -m gpu > gpurun_out/packed_tests_v2.log 2>&1; echo "v2 rc=$?"; tail -3 gpurun_out/packed_tests_v2.log
python tools/k1_ab.py mandelmesh2048:0,1,2 torus2048:0,1,2 martin_cube1024:0,2 p_key1024:0,2 > gpurun_out/k1_ab.jsonl 2> gpurun_out/k1_ab.err; cat gpurun_out/k1_ab.jsonl; tail -2 gpurun_out/k1_ab.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_packed.json 2> gpurun_out/bench_packed.err; tail -c 400 gpurun_out/bench_packed.json
timeout 240 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_all.log 2>&1; echo "all rc=$?"; tail -4 gpurun_out/gpu_tests_all.log
