for r in 1 2; do
  for w in mandelmesh2048 torus2048; do
    S2M_K1_ROWS=$r python bench.py --no-cpu-baseline --workload $w | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('rows=$r', '$w', round(d['e2e']['value'],1), round(d['ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['kernels'].items() if isinstance(v,dict)})"
  done
done
S2M_K1_ROWS=2 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -3
