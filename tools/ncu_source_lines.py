#!/usr/bin/env python3
"""Executed warp instructions per SOURCE LINE of a kernel from an ncu --set full --import-source on capture.

    python tools/ncu_source_lines.py gpurun_out/x.ncu-rep [top N] [kernel index]

NVRTC headers are in-memory, so ncu cannot show their text; the line numbers are resolved against sdf2mesh_b200/csrc/.
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, key = None, None
    agg, opagg, smp = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
    total = 0
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = os.path.basename(r[1]); continue
        if r[0] in ("Function Name", "Line No", "Kernel Name"):
            continue
        if r[0] != "":
            key = (cur, int(r[0])); continue
        try:
            ie = int(r[7])
        except (ValueError, IndexError):
            continue
        agg[key] += ie; total += ie
        smp[key] += int(r[6]) if r[6].isdigit() else 0
        toks = r[3].split()
        op = toks[0] if not toks[0].startswith('@') else toks[1]
        opagg[key][op.split('.')[0]] += ie
    files = {}

    def src(f, l):
        p = os.path.join(ROOT, "sdf2mesh_b200", "csrc", f)
        if f not in files:
            files[f] = open(p).read().splitlines() if os.path.exists(p) else None
        return files[f][l - 1].strip()[:120] if files[f] and l <= len(files[f]) else "?"
    byfile = collections.Counter()
    for (f, l), n in agg.items():
        byfile[f] += n
    print("total warp instructions:", total)
    print("by file (%):", {f: round(n / total * 100, 1) for f, n in byfile.most_common()})
    ops = collections.Counter()
    for k in opagg:
        ops.update(opagg[k])
    print("by opcode (%):", {o: round(n / total * 100, 1) for o, n in ops.most_common(24)})
    ssum = sum(smp.values()) or 1
    for (f, l), n in agg.most_common(top):
        print(f"{n / total * 100:5.2f}% (samples {smp[(f, l)] / ssum * 100:5.2f}%) {f}:{l}  {dict(opagg[(f, l)].most_common(4))}  | {src(f, l)}")


if __name__ == "__main__":
    main()
