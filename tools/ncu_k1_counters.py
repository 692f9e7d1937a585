#!/usr/bin/env python3
"""K1's per-step counters from an ncu metrics pass -> profiles/k1_counters.json (read by bench.py for the roofline).

    ncu --metrics <METRICS below> --clock-control none -k regex:s2m_k1_slab --csv --log-file gpurun_out/k1cnt_<wl>.csv \\
        python bench.py --workload <wl> --steps 1 --warmup 1 --no-verify --no-cpu-baseline --no-other-workloads

    python tools/ncu_k1_counters.py <wl>=gpurun_out/k1cnt_<wl>.csv[:<K1 launches per step>] ...

That bench command runs 5 meshing steps (warm-up, timed, one with event spans, two serialised); the launches of the
SECOND step are summed.  <K1 launches per step> = z-chunks of a pipelined step = gpu_launches / (5 * steps) of a
bench line of the same workload.
Counters are properties of the instruction stream and the grid, so bench.py divides them by the K1 time it measures live.
"""
import csv
import json
import sys

METRICS = ("gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,"
           "sm__sass_thread_inst_executed_op_fadd_pred_on.sum,sm__sass_thread_inst_executed_op_fmul_pred_on.sum,sm__sass_thread_inst_executed_op_ffma_pred_on.sum,"
           "sm__sass_thread_inst_executed_op_fadd2_pred_on.sum,sm__sass_thread_inst_executed_op_fmul2_pred_on.sum,sm__sass_thread_inst_executed_op_ffma2_pred_on.sum,"
           "sm__sass_thread_inst_executed_op_fp32_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum")
FLOPS = {"fadd": 1, "fmul": 1, "ffma": 2, "fadd2": 2, "fmul2": 2, "ffma2": 4}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "inst": 1.0, "": 1.0}


def parse(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = {}
    for r in rows[1:]:
        if "s2m_k1_slab" not in r[ix["Kernel Name"]]:
            continue
        d = launches.setdefault(int(r[ix["ID"]]), {"grid": r[ix["Grid Size"]]})
        d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * UNIT.get(r[ix["Metric Unit"]], 1.0)
    return [launches[k] for k in sorted(launches)]


def main():
    out = {}
    for spec in sys.argv[1:]:
        wl, path = spec.split("=", 1)
        path, _, per_s = path.partition(":")
        ls = parse(path)
        if per_s:
            per = int(per_s)      # z-chunks (= K1 launches) of a pipelined step: gpu_launches / (5 * steps) of the bench line
        else:                     # ... or found here: the first three steps are pipelined alike (their grid sequence repeats
            grids = [l["grid"] for l in ls]   # three times), the two serialised steps that follow are alike too
            ok = [q for q in range(1, len(grids) // 3 + 1)
                  if grids[:q] == grids[q:2 * q] == grids[2 * q:3 * q] and len(grids) > 3 * q and (len(grids) - 3 * q) % 2 == 0]
            per = max(ok) if ok else max(1, len(grids) // 5)
        step = ls[per:2 * per]    # the second step (the first pipelined step after the warm-up)
        tot = lambda m: sum(l.get(m, 0.0) for l in step)
        op = lambda o: tot(f"sm__sass_thread_inst_executed_op_{o}_pred_on.sum")
        flops = sum(op(o) * f for o, f in FLOPS.items())
        heavy = max(step, key=lambda l: l.get("gpu__time_duration.sum", 0.0))
        out[wl] = {
            "launches_per_step": per, "grid_of_the_longest_launch": heavy["grid"],
            "k1_ns_under_ncu_per_step": tot("gpu__time_duration.sum"),
            "warp_inst_per_step": tot("smsp__inst_executed.sum"), "thread_inst_per_step": tot("smsp__thread_inst_executed.sum"),
            "fp32_flops_per_step": flops, "fp32_thread_inst_per_step": tot("sm__sass_thread_inst_executed_op_fp32_pred_on.sum"),
            "op_thread_inst_per_step": {o: op(o) for o in FLOPS},
            "dram_bytes_per_step": tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum"),
            "dram_bytes_per_launch": heavy.get("dram__bytes_read.sum", 0.0) + heavy.get("dram__bytes_write.sum", 0.0),
            "source": f"profiles/{path.split('/')[-1]} (ncu --metrics, second of four steps)",
        }
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
