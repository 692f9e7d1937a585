set -x
B="python bench.py --steps 1 --warmup 1 --no-verify --no-cpu-baseline --no-other-workloads"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2m_k1_slab -s 17 -c 1 -o gpurun_out/r02_k1_mandel_opt $B > /dev/null 2>&1
S2M_K1_OPTIMISTIC=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2m_k1_slab -s 17 -c 1 -o gpurun_out/r02_k1_mandel_exact $B > /dev/null 2>&1
ls -la gpurun_out/r02_k1_mandel_opt.ncu-rep gpurun_out/r02_k1_mandel_exact.ncu-rep
