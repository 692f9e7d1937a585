# One B200: K1's per-step ncu counters (tools/ncu_k1_counters.py METRICS) for the workloads named on the command line
# usage: bash tools/gpu_k1cnt.sh mandelmesh2048 [torus2048 ...]
set -x
B="python bench.py --steps 1 --warmup 1 --no-verify --no-cpu-baseline --no-other-workloads"
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_k1_counters as n; print(n.METRICS)")
for wl in "$@"; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:s2m_k1_slab --csv --log-file gpurun_out/k1cnt_$wl.csv $B --workload $wl > /dev/null 2> gpurun_out/k1cnt_$wl.err
  wc -l gpurun_out/k1cnt_$wl.csv
done
