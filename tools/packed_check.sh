# one-GPU validation of the packed K1: A/B timings, parity, the whole GPU suite, the default bench line, launch list
set -x
mkdir -p gpurun_out
python tools/k1_ab.py mandelmesh2048:0,d mandelmesh2048:2:S2M_K1_ROWS=1,S2M_K1_MINBLOCKS=6 torus2048:0,d p_key1024:d martin_cube1024:d > gpurun_out/k1_ab_final.jsonl 2> gpurun_out/k1_ab.err; cat gpurun_out/k1_ab_final.jsonl; tail -2 gpurun_out/k1_ab.err
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_all.log 2>&1; echo "all rc=$?"; tail -3 gpurun_out/gpu_tests_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_packed.json 2> gpurun_out/bench_packed.err; tail -c 300 gpurun_out/bench_packed.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_packed.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
