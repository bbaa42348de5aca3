#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.
usage: summarize_profiles.py <tag> <launches.csv> [<kernel.ncu-rep> <kernel-label>]"""
import collections
import csv
import re
import subprocess
import sys

tag, launches = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn])
    f = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[mv].replace(",", "")) * f
tot = sum(a[1] for a in agg.values())
with open(f"profiles/{tag}_launches.md", "w") as out:
    out.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none`\n\n")
    out.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
    out.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"| `{k.strip()}` | {c} | {t:.3f} | {100 * t / tot:.1f}% |\n")
    out.write(f"\ntotal {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
if len(sys.argv) > 4:
    rep, label = sys.argv[3], sys.argv[4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h = r[0]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
            "launch__grid_size", "launch__block_size"]
    with open(f"profiles/{tag}_{label}_ncu.md", "w") as out:
        out.write(f"# ncu --set full ({tag}): {label}\n\n| metric | unit | " +
                  " | ".join(f"launch {i}" for i in range(len(r) - 2)) + " |\n|---|---|" +
                  "---|" * (len(r) - 2) + "\n")
        for w in want:
            if w in h:
                i = h.index(w)
                out.write(f"| {w} | {r[1][i]} | " + " | ".join(x[i] for x in r[2:]) + " |\n")
