# One B200: --set full captures of a light (all corners far outside) and the heaviest K1 launch of mandelmesh 2048^3
set -x
B="python bench.py --steps 1 --warmup 1 --no-verify --no-cpu-baseline --no-other-workloads"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2m_k1_slab -s 13 -c 1 -o gpurun_out/r02_k1_mandel_light $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2m_k1_slab -s 17 -c 1 -o gpurun_out/r02_k1_mandel_heavy $B > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
