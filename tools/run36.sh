for pad in 0 20000 40000; do
  S2M_K4_SMEM_PAD=$pad python bench.py --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pad=$pad', round(d['e2e']['value'],1), round(d['ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['kernels'].items() if isinstance(v,dict)})"
done
for w in mandelmesh1024 mandelmesh512 torus2048 p_key1024; do
  python bench.py --no-cpu-baseline --workload $w | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['e2e']['value'],1), round(d['ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['kernels'].items() if isinstance(v,dict)})"
done
