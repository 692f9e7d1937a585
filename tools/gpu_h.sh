# plane carry-over: parity, bench, trace at a small slab (N=8-like), cold CLI timings, sanitizer, smoke, K1 counters for the new launch shapes
set -x
timeout 1800 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_h.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_h.err
S2M_CARRY_PLANES=0 timeout 600 python bench.py --no-cpu-baseline --no-other-workloads --no-verify > gpurun_out/bench_h_nocarry.json 2>> gpurun_out/bench_h.err
timeout 300 python tools/host_side_timings.py cli > gpurun_out/cli_timings_r02.jsonl 2>&1; cat gpurun_out/cli_timings_r02.jsonl | cut -c1-900
python - <<'PY' 2> gpurun_out/trace_h.txt
import os, sys
sys.path.insert(0, ".")
os.environ["S2M_TRACE"] = "1"
import sdf2mesh_b200 as s2m
ctx = s2m.Context(0)
m = s2m.Sdf3DShader.from_glsl_fragment_shader("examples/mandelmesh.frag", "sdf").create_shader_module(ctx)
p, _ = s2m.params_from_cli(2048, 5.0, flags=s2m.MESH_QUADS_U32 | s2m.MESH_RELATIVE_QUADS)
p.z_begin, p.z_end = 872, 1018     # rank 3 of 8
for i in range(4):
    r = s2m.mesh_begin(ctx, m, p); r.finish(1000); r.free()
PY
tail -80 gpurun_out/trace_h.txt
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/sanitize_r02_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitize_r02_$tool.log
done
