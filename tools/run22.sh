set -x
python -m pytest tests/test_aux_gpu.py -x -q -m gpu -k "parts or cli" 2>&1 | tail -5
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/sanitize_$tool.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:s2m_k1_slab -s 12 -c 1 -o gpurun_out/k1_2048_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1.log 2>&1
echo ncu rc=$?; tail -3 gpurun_out/ncu_k1.log
ls -la gpurun_out/*.ncu-rep
