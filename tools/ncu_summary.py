#!/usr/bin/env python
"""Builds profiles/ncu_summary_rNN.md from the captures a gpurun call left in gpurun_out/ (launch list CSV + .ncu-rep files).
Usage: python tools/ncu_summary.py <round-tag> <launches.csv> <rep> [<rep> ...]"""
import collections
import csv
import json
import subprocess
import sys

tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
out = ["# Round-%s ncu evidence (B200, sm_100a, driver 580, CUDA 12.9)" % ("2" if "r02" in tag else "1"), "",
       "Command (run plain first, then under ncu): `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph`",
       "Launch list: `profiles/launches_%s.csv` (`ncu --metrics gpu__time_duration.sum --clock-control none`).  Per-launch times are "
       "cold-cache and serialised (ncu flushes caches between replays): compare SHARES, not absolutes." % tag, ""]
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
seq = [(r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("bode::", ""), float(r[ix["Metric Value"]]) / 1e3)
       for r in rows[1:] if r[ix["Metric Name"]] == "gpu__time_duration.sum"]
# one step = from one npde grad launch to the next; take the LAST complete step of the eager loop
idx = [i for i, (n, _) in enumerate(seq) if "npde_pair_grad" in n or "npde_grad_kernel" in n]
step = []
for a, b in zip(idx[:-1], idx[1:]):
    names = [n for n, _ in seq[a:b]]
    if any("phi2" in n or "phi_tc" in n or "phi_partial" in n for n in names) and any("gram" in n or "sqdist" in n for n in names):
        step = seq[a:b]
agg = collections.OrderedDict()
for n, us in step:
    if "Fill" in n or "elementwise" in n:
        continue                      # the L2 flush memset between steps (outside the timed region)
    c = agg.setdefault(n, [0, 0.0])
    c[0] += 1
    c[1] += us
tot = sum(v[1] for v in agg.values())
out += ["## Share of one SVGD sampler step (c3: P=4096, 39 rk4 steps), from the launch list", "",
        "| kernel | launches/step | us/step | share |", "|---|---|---|---|"]
for n, (c, us) in agg.items():
    out.append("| %s | %d | %.1f | %.1f%% |" % (n, c, us, 100 * us / tot))
out += ["| total | %d | %.1f | 100%% |" % (sum(v[0] for v in agg.values()), tot), ""]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max"]
traffic = {}
seen = set()          # a kernel is reported from the FIRST capture that holds it: list the newest capture first
for rep in reps:
    csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(csvtxt.splitlines()))
    h, units = rr[0], rr[1]
    for r in rr[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("bode::", "")
        if name in seen:
            continue
        seen.add(name)
        out += ["### %s  (%s, `ncu --set full --clock-control none --import-source on`)" % (name, rep.split("/")[-1])]
        for k in keys:
            if k in h:
                out.append("- %s = %s %s" % (k, r[h.index(k)], units[h.index(k)]))
        st = []
        for i, hh in enumerate(h):
            if "issue_stalled" in hh and hh.endswith("per_issue_active.ratio") and "not_issued" not in hh:
                try:
                    st.append((float(r[i]), hh.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        st.sort(reverse=True)
        out.append("- top stall reasons (warps per issue): " + ", ".join("%s=%.2f" % (n, v) for v, n in st[:6]))
        out.append("")
        def num(k):
            v = float(r[h.index(k)])
            u = units[h.index(k)].lower()
            return v * (1e9 if u.startswith("gb") else 1e6 if u.startswith("mb") else 1e3 if u.startswith("kb") else 1.0)
        traffic[name] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
open("profiles/ncu_summary_%s.md" % tag, "w").write("\n".join(out) + "\n")
json.dump({"source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch", "bytes_per_launch": traffic},
          open("profiles/traffic_%s.json" % tag, "w"), indent=1)
print("\n".join(out[:30]))
