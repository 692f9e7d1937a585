python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for e in 0 1; do
  if [ $e = 1 ]; then export S2M_NO_CHUNK_OVERLAP=1; fi
  for w in mandelmesh2048 torus2048; do
    python bench.py --no-cpu-baseline --workload $w | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no_overlap=$e', '$w', round(d['e2e']['value'],1), round(d['ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['kernels'].items() if isinstance(v,dict)}, d['kernels']['device_total_ms'])"
  done
done
unset S2M_NO_CHUNK_OVERLAP
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/sanitize2_$tool.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/sanitize2_$tool.log
done
