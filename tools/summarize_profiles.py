#!/usr/bin/env python
"""Turns the ncu artefacts in gpurun_out/ into the small text summaries committed under profiles/."""
import csv
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__cluster_size"]


def launch_list(name, per_step):
    path = os.path.join(G, name)
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    ks = [(r["Kernel Name"], r["Grid Size"], float(r["Metric Value"].replace(",", "")) / 1000.0) for r in rows]
    last = ks[-per_step:]
    tot = sum(k[2] for k in last)
    out = ["# %s: last step's launches (gpu__time_duration, us; cold-cache, serialised under ncu --clock-control none)" % name,
           "# share = duration / sum of the step's launches (%.1f us)" % tot]
    agg = OrderedDict()
    for n, g, d in last:
        out.append("%-100s grid %-14s %8.2f us  %5.1f %%" % (n[:100], g, d, 100 * d / tot))
        key = n.split("(")[0][:70]
        agg[key] = agg.get(key, 0.0) + d
    out.append("# by kernel:")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        out.append("#   %-72s %8.2f us  %5.1f %%" % (k, v, 100 * v / tot))
    open(os.path.join(P, name.replace(".csv", ".txt")), "w").write("\n".join(out) + "\n")
    print("\n".join(out[-8:]))


def raw_page(rep):
    """raw metric page of a report: the exported <name>_raw.csv (tools/run_profiles.sh) or, if the report itself is here, ncu -i"""
    path = os.path.join(G, rep)
    csv_path = path.replace(".ncu-rep", "_raw.csv")
    if os.path.exists(csv_path):
        return open(csv_path).read()
    return subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout


def full_report(rep, out_name):
    txt = raw_page(rep)
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    out = ["# %s (ncu --set full --clock-control none), one line block per profiled launch" % rep]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append("kernel: %s  grid %s block %s" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        for k in KEYS:
            for h in hdr:
                if h == k:
                    out.append("    %-80s %s" % (k, d[h]))
    open(os.path.join(P, out_name), "w").write("\n".join(out) + "\n")
    print(out_name, "written,", len(rows) - 2, "launches")


def traffic(rep, kernel_substr):
    """dram read+write bytes per launch of the longest launch of the named kernel in a --set full report."""
    txt = raw_page(rep)
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    best = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        subs = kernel_substr if isinstance(kernel_substr, (tuple, list)) else (kernel_substr,)
        if not any(k in d.get("Kernel Name", "") for k in subs):
            continue
        t = float(d["gpu__time_duration.sum"].replace(",", ""))
        if best is None or t > best[0]:
            u = dict(zip(hdr, units))
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(d["dram__bytes_read.sum"].replace(",", "")) * scale[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"].replace(",", "")) * scale[u["dram__bytes_write.sum"]]
            best = (t, rd + wr)
    return None if best is None else best[1]


if __name__ == "__main__":
    import json
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01n"
    have = lambda n: os.path.exists(os.path.join(G, n)) or os.path.exists(os.path.join(G, n.replace(".ncu-rep", "_raw.csv")))
    tpath = os.path.join(P, "traffic.json")
    tr = json.load(open(tpath)) if os.path.exists(tpath) else {}
    # a configuration whose kernels did not change since an earlier tag keeps that tag's files (and its traffic entry)
    per = {"single": 16, "64seeds": 19, "8seeds": 19}
    if have(tag + "_launches_per_step.txt"):
        w = open(os.path.join(G, tag + "_launches_per_step.txt")).read().split()
        per.update({w[i]: int(w[i + 1]) for i in range(0, len(w), 2)})
    if have(tag + "_launches_single_fp32.csv"):
        launch_list(tag + "_launches_single_fp32.csv", per["single"])
    if have(tag + "_launches_64seeds_tf32.csv"):
        launch_list(tag + "_launches_64seeds_tf32.csv", per["64seeds"])
    if have(tag + "_launches_8seeds_tf32.csv"):
        launch_list(tag + "_launches_8seeds_tf32.csv", per["8seeds"])
    if have(tag + "_8seeds_tf32.ncu-rep"):
        full_report(tag + "_8seeds_tf32.ncu-rep", tag + "_8seeds_tf32_full.txt")
        tr["8:tf32"] = traffic(tag + "_8seeds_tf32.ncu-rep", ("gemm_ws_kernel", "gemm_ws2_kernel", "gemm_chain_kernel"))
    if have(tag + "_single_fp32.ncu-rep"):
        full_report(tag + "_single_fp32.ncu-rep", tag + "_single_fp32_full.txt")
        tr["1:fp32"] = traffic(tag + "_single_fp32.ncu-rep", ("gemm_sk_kernel", "gemm_fwd2_kernel"))
    if have(tag + "_64seeds_tf32.ncu-rep"):
        full_report(tag + "_64seeds_tf32.ncu-rep", tag + "_64seeds_tf32_full.txt")
        tr["64:tf32"] = traffic(tag + "_64seeds_tf32.ncu-rep", ("gemm_ws_kernel", "gemm_ws2_kernel"))
    # round-2 extras: the generic glue kernels before their many-seed rewrite, and the GEMM stages alone
    if have("r02c_glue64.ncu-rep"):
        full_report("r02c_glue64.ncu-rep", "r02_glue64_before.txt")
    if have("r02e_gemm64.ncu-rep"):
        full_report("r02e_gemm64.ncu-rep", "r02_gemm64_full.txt")
    json.dump(tr, open(tpath, "w"), indent=1)
    print(tr)
