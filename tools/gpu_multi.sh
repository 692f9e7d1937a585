# N GPUs (gpurun --gpus N -- bash tools/gpu_multi.sh): the s2m_multi / NCCL tests, then the 2048^3 bench one process per GPU (torchrun), from one
# process through s2m_multi_mesh_run (NCCL count exchange) and with the host-memory exchange
set -x
N=$(nvidia-smi -L | wc -l)
# S2M_MULTI_REDUCED=1 (8 GPUs cost 8x the box time): no tests, no host-exchange run, the one-process run without the other workloads
[ -n "$S2M_MULTI_REDUCED" ] || timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_aux_gpu.py -q -m gpu -k "multi or nccl or cli" --timeout 600 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_r2_n$N.json 2> gpurun_out/scale_r2_n$N.err; echo "ranks rc=$?"; tail -3 gpurun_out/scale_r2_n$N.err
timeout 600 python bench.py --gpus $N --driver capi-multi --steps 20 --warmup 5 ${S2M_MULTI_REDUCED:+--no-other-workloads} > gpurun_out/multi_r2_n$N.json 2> gpurun_out/multi_r2_n$N.err; echo "capi-multi rc=$?"; tail -3 gpurun_out/multi_r2_n$N.err
[ -n "$S2M_MULTI_REDUCED" ] || timeout 600 python bench.py --gpus $N --driver capi-multi --no-nccl --steps 20 --warmup 5 --no-other-workloads > gpurun_out/multi_nonccl_r2_n$N.json 2>> gpurun_out/multi_r2_n$N.err; echo "capi-multi no-nccl rc=$?"
python - <<PY
import json
for f in ("scale_r2_n$N","multi_r2_n$N","multi_nonccl_r2_n$N"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, "value=%.1f e2e=%.1f ms=%.3f parity=%s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity_check"] and d["parity_check"]["ok"]))
        for r in d["per_rank"] or []: print("   ", {k:r[k] for k in r if k in ("rank","device","slices","begin_ms","allgather_wait_ms","exchange_ms","finish_ms","k1_slab_ms","device_ms")})
        for o in d.get("other_workloads") or []: print("   other", o["workload"], "value=%.1f e2e=%.1f ms=%.3f"%(o["value"], o["e2e"]["value"], o["ms_per_step"]), o["parity_check"] and o["parity_check"]["ok"])
    except Exception as e: print(f, "ERR", e)
PY
