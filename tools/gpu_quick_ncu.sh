# One B200: tools/gpu_quick.sh, then a --set full capture of a K1 launch outside the fractal (mandelmesh 2048^3, second step, chunk 1)
bash tools/gpu_quick.sh "$@"
B="python bench.py --steps 1 --warmup 1 --no-verify --no-cpu-baseline --no-other-workloads"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2m_k1_slab -s 13 -c 1 -f -o gpurun_out/r02_k1_mandel_light $B > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
