# first GPU pass of round 2: whole -m gpu suite on the restructured pipeline, then the benches
set -x
nvidia-smi -L
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_a.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_a.json; tail -5 gpurun_out/bench_a.err
timeout 300 python bench.py --no-cpu-baseline --workload torus2048 > gpurun_out/bench_a_torus.json 2>> gpurun_out/bench_a.err; tail -c 600 gpurun_out/bench_a_torus.json
