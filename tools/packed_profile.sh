# FFMA vs FFMA2 throughput, then ncu --set full of one heavy K1 launch, packed and scalar (same chunk)
set -x
mkdir -p gpurun_out
./tools/ubench/ffma2_rate > gpurun_out/ffma2_rate.jsonl 2>&1; cat gpurun_out/ffma2_rate.jsonl
ncu --set full --clock-control none --import-source on -k regex:s2m_k1_slab -s 12 -c 1 -f -o gpurun_out/k1_packed python tools/k1_ab.py mandelmesh2048:1 > gpurun_out/ncu_packed.log 2>&1; tail -2 gpurun_out/ncu_packed.log
ncu --set full --clock-control none -k regex:s2m_k1_slab -s 12 -c 1 -f -o gpurun_out/k1_scalar python tools/k1_ab.py mandelmesh2048:0 > gpurun_out/ncu_scalar.log 2>&1; tail -2 gpurun_out/ncu_scalar.log
ls -la gpurun_out/*.ncu-rep
