# round-end style validation on one B200: tests, smoke, both bench arms, sanitizer, ncu launch list + full captures
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err; tail -c 200 gpurun_out/bench_final_reference.json
python bench.py --workload mandelmesh4096 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final_4096.json 2>> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final_4096.json
python tools/host_side_timings.py jit writer cli > gpurun_out/host_side_timings.jsonl 2>&1; cat gpurun_out/host_side_timings.jsonl
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/sanitize_final_$tool.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/sanitize_final_$tool.log
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"s2m_k1_slab|k2_classify|s2m_k4_vertices|k3_compact|k4_quads" -s 60 -c 5 -o gpurun_out/final_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/final_full.ncu-rep gpurun_out/launches_final.csv
