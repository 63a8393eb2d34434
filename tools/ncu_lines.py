#!/usr/bin/env python
"""Per-source-line totals (instructions executed, stall samples, shared-memory wavefronts) from an
`ncu --set full --import-source on` report of a kernel compiled with -lineinfo.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n] [--by-samples]
"""
import csv, io, subprocess, sys, collections

by_samples = "--by-samples" in sys.argv
argv = [a for a in sys.argv if a != "--by-samples"]
rep = argv[1]
top = int(argv[2]) if len(argv) > 2 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# layout: blocks of [cuda line header rows ...]; find the header of the correlated table
hdr_i = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0, 0, 0, ""])
fname = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) < len(hdr) or not r[0].isdigit():
        continue                                   # SASS rows have an empty "Line No"

    def num(k):
        try:
            return int(float(r[col[k]]))
        except Exception:
            return 0
    a = agg[(fname, int(r[0]))]
    a[0] += num("Instructions Executed"); a[1] += num("# Samples")
    a[2] += num("L1 Wavefronts Shared"); a[3] += num("L1 Wavefronts Shared Excessive")
    a[4] = r[1]
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print("total inst %d, samples %d" % (tot, tots))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_samples else 0])[:top]:
    print("%-22s inst %5.1f%%  samples %5.1f%%  smem_wf %9d excess %9d  %s" % ("%s:%d" % k, 100.0 * a[0] / max(tot, 1), 100.0 * a[1] / max(tots, 1), a[2], a[3], a[4].strip()[:100]))
