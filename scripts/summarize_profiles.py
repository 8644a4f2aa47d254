"""Turns the ncu outputs that scripts/capture_profiles.sh left in gpurun_out/ into the committed
summaries under profiles/ (run here, no GPU needed).  usage: python scripts/summarize_profiles.py r01"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
os.makedirs(P, exist_ok=True)


def launch_list():
    src = os.path.join(G, "launches_bench.csv")
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            agg.setdefault(row["Kernel Name"], []).append(float(row["Metric Value"]))
        except Exception:
            pass
    tot = sum(sum(v) for v in agg.values())
    out = [f"# ncu launch list — `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu` ({tag})", "",
           "(the command makes two cold PtAP calls, 5 steps of numeric PtAP + M^T b + CG, and the per-phase timing loops)", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -c 4000`; per-launch times are cold-cache and",
           "serialised: read the SHARES.  N_b = 184 (50.2 M foreground dofs), 1 x B200.", "",
           f"total kernel time in the capture: {tot / 1e6:.1f} ms over {sum(len(v) for v in agg.values())} launches", "",
           "| kernel | launches | total ms | share | mean us |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        name = k.replace("iife::", "").split("(")[0]
        out.append(f"| `{name}` | {len(v)} | {sum(v) / 1e6:.2f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v) / 1e3:.1f} |")
    open(os.path.join(P, f"{tag}_launch_list.md"), "w").write("\n".join(out) + "\n")
    with open(os.path.join(P, f"{tag}_launches_bench.csv"), "w") as f:
        f.writelines(lines)


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic"]


def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return []
    h, u = rows[0], rows[1]
    res = []
    for v in rows[2:]:
        d = {"Kernel Name": v[h.index("Kernel Name")]}
        for k in KEYS:
            if k in h:
                d[k] = (v[h.index(k)], u[h.index(k)])
        res.append(d)
    return res


def stalls(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    parts = txt.split('"Kernel Name"')
    if len(parts) < 2:
        return {}, 0
    rows = list(csv.reader(('"Kernel Name"' + parts[1]).splitlines()))
    h, data = rows[1], [r for r in rows[2:] if len(r) > 10]
    ix = {n: i for i, n in enumerate(h)}
    names = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in names}
    inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
    return {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v}, inst


def kernel_report(rep_name, title, fname, note=""):
    rep = os.path.join(G, rep_name)
    if not os.path.exists(rep):
        return None
    ms = raw_metrics(rep)
    st, inst = stalls(rep)
    out = [f"# {title} ({tag})", "", f"`ncu --set full --clock-control none` on the bench command, N_b = 184, 1 x B200.  {note}", ""]
    for d in ms:
        out.append(f"## `{d['Kernel Name'].split('(')[0]}`")
        out.append("")
        out.append("| metric | value | unit |")
        out.append("|---|---:|---|")
        for k in KEYS:
            if k in d:
                out.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
        out.append("")
    if st:
        tot = sum(st.values())
        out.append("## warp-stall samples (first captured launch)")
        out.append("")
        out.append(", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in list(st.items())[:8]))
        out.append("")
        out.append(f"warp instructions executed: {inst}")
    open(os.path.join(P, fname), "w").write("\n".join(out) + "\n")
    return ms


launch_list()
sell = kernel_report("prof_spmv_sell.ncu-rep", "SELL-32 SpMV of the CG iteration (k_spmv_sell)", f"{tag}_spmv_sell.md")
kernel_report("prof_ptap_numeric.ncu-rep", "Numeric PtAP (template kernel k_ptap_numeric_tpl, two rows per warp)", f"{tag}_ptap_numeric.md",
              note="Per output row (6 331 625 rows): divide smsp__inst_executed.sum and the byte counts by the row count.")
kernel_report("prof_ptap_symbolic.ncu-rep", "Symbolic PtAP (count and fill pass of k_ptap_symbolic, cold path)", f"{tag}_ptap_symbolic.md")
kernel_report("prof_cg_vec.ncu-rep", "CG vector kernels", f"{tag}_cg_vector_kernels.md")
if sell:
    # the DOT variant is the kernel inside the iteration
    d = [m for m in sell if "k_spmv_sell" in m["Kernel Name"]][0]
    rd = float(d["dram__bytes_read.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"][1]]
    wr = float(d["dram__bytes_write.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][1]]
    json.dump({"n_bg_cells": 184, "kernel": d["Kernel Name"].split("(")[0], "spmv_dot_dram_bytes_per_launch": rd + wr,
               "source": f"profiles/{tag}_spmv_sell.md"}, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
for src, dst in (("configs_1_4.md", f"{tag}_configs_1_4.md"), ("robustness.md", f"{tag}_robustness.md"), ("phase184.log", f"{tag}_phase184.log")):
    if os.path.exists(os.path.join(G, src)):
        open(os.path.join(P, dst), "w").write(open(os.path.join(G, src)).read())
print(open(os.path.join(P, f"{tag}_launch_list.md")).read()[:1800])
