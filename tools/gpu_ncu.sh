# One B200: ncu evidence -- per-step K1 counters for every workload (-> tools/ncu_k1_counters.py -> profiles/k1_counters.json),
# the launch list of the bench command and --set full captures of K1 (per example SDF) and of K2 / K3 / K4a / K4b
set -x
B="python bench.py --steps 1 --warmup 1 --no-verify --no-cpu-baseline --no-other-workloads"
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_k1_counters as n; print(n.METRICS)")
for wl in mandelmesh2048 torus2048 martin_cube512 p_key1024 p_key1024_b20 torus128; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:s2m_k1_slab --csv --log-file gpurun_out/k1cnt_$wl.csv $B --workload $wl > /dev/null 2> gpurun_out/k1cnt_$wl.err
  wc -l gpurun_out/k1cnt_$wl.csv
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_mandelmesh2048.csv python bench.py --steps 2 --warmup 1 --no-verify --no-cpu-baseline --no-other-workloads > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2m_k1_slab -s 13 -c 1 -f -o gpurun_out/r02_k1_mandel $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2m_k1_slab -s 12 -c 1 -f -o gpurun_out/r02_k1_torus $B --workload torus2048 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2m_k1_slab -s 2 -c 1 -f -o gpurun_out/r02_k1_martin $B --workload martin_cube512 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2m_k1_slab -s 10 -c 1 -f -o gpurun_out/r02_k1_pkey $B --workload p_key1024 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2_classify|k3_compact|s2m_k4_vertices|k4_quads" -s 52 -c 4 -f -o gpurun_out/r02_k234 $B > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
