#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full) into the markdown summary committed under profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" "command line that produced it" > profiles/xxx.md
"""
import csv
import io
import subprocess
import sys

KEYS = [('gpu__time_duration.sum', 'duration'), ('smsp__inst_executed.sum', 'warp instructions'),
        ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active threads / instruction'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
        ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe active %'),
        ('sm__inst_executed_pipe_xu.sum', 'XU (MUFU) pipe instructions'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
        ('launch__registers_per_thread', 'registers/thread'), ('dram__bytes_read.sum', 'DRAM read'),
        ('dram__bytes_write.sum', 'DRAM write'), ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
        ('lts__t_bytes.sum', 'L2 bytes'), ('l1tex__t_bytes.sum', 'L1 bytes')]


def main():
    rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
    out = [f"# {title}", "", f"Command (under gpurun): `{cmd}`", "",
           "Times under ncu are cold-cache and serialised: compare shares, not absolutes.", ""]
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0]
        out.append(f"## {name}   grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        out += ["", "| metric | value |", "|---|---|"]
        for k, label in KEYS:
            if k in idx:
                out.append(f"| {label} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
        st = sorted(((float(r[idx[h]] or 0), h) for h in stall), reverse=True)[:5]
        out.append("| top stall reasons (warps stalled per issue-active cycle) | " +
                   "; ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, h in st) + " |")
        out.append("")
    print("\n".join(out))


if __name__ == "__main__":
    main()
