#!/usr/bin/env python
"""Top CUDA source lines of one kernel in an .ncu-rep by executed warp instructions / stall samples.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep <substring of the kernel's demangled name> [top n]   (all matching launches are summed)
"""
import csv
import subprocess
import sys
from collections import defaultdict

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout        # rx: substring of the demangled function name
rows = list(csv.reader(out.splitlines()))
agg, hdr, fpath, name, keep = {}, None, None, "", False
tot_i = tot_s = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        keep = rx in r[1]
        if keep:
            name = r[1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or not keep or len(r) != len(hdr) or not r[0].strip().isdigit():
        continue                      # SASS rows carry an empty line number: the per-line rows already aggregate them
    try:
        n = int(r[ix["Instructions Executed"]] or 0); sm = int(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    key = (fpath, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1][:100]])
    a[0] += n; a[1] += sm
    tot_i += n; tot_s += sm
print("#", name)
print("# total warp instructions %d, stall samples %d" % (tot_i, tot_s))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1 if "--by-samples" in sys.argv else 0])[:top]:
    print("%5.1f%% inst %5.1f%% samples  %s:%d  %s" % (100.0 * v[0] / max(tot_i, 1), 100.0 * v[1] / max(tot_s, 1), k[0], k[1], v[2]))
