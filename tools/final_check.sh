# round-end style validation: tools/gpu_validate.sh (tests, smoke, benches, CLI timings, sanitizer) and tools/gpu_ncu.sh
# (K1 counters per workload, launch list, --set full captures) on one B200; tools/gpu_multi.sh on 2 / 8 GPUs
bash tools/gpu_validate.sh
bash tools/gpu_ncu.sh
