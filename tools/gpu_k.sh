set -x
timeout 900 python tools/k1_ab.py mandelmesh2048:d mandelmesh2048:d:S2M_EXP_OPT=1 > gpurun_out/k1_exp.jsonl 2> gpurun_out/k1_exp.err; cat gpurun_out/k1_exp.jsonl | cut -c1-400; tail -3 gpurun_out/k1_exp.err
