# One B200, end of a round: the whole -m gpu suite, smoke(), and one default bench line (no CPU baseline) at HEAD
set -x
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_final.json") if l.startswith("{")][-1])
print("value %.1f e2e %.1f ms %.3f parity %s launches %d roofline %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity_check"]["ok"], d["gpu_launches"], d["roofline"]["frac"]))
for o in d["other_workloads"]: print("  ", o["workload"], "%.3f ms"%o["ms_per_step"], (o.get("parity_check") or {}).get("ok"))
PY
