# One B200 (gpurun -- 'bash tools/gpu_validate.sh'): the whole -m gpu suite, smoke(), both bench arms, the cold / warm CLI
# timings and compute-sanitizer over every kernel path.  Outputs under gpurun_out/.
set -x
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2>> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_reference_arm.json
timeout 300 python tools/host_side_timings.py cli > gpurun_out/cli_timings.jsonl 2>&1; cut -c1-600 gpurun_out/cli_timings.jsonl
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitize_$tool.log
done
