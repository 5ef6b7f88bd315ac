#!/usr/bin/env python
"""Summarise ncu outputs into profiles/ (tracked).
  python tools/summarize_ncu.py launches gpurun_out/launches_rX.csv profiles/rX_launches.md
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/rX_kernel.md
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    by_grid = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if not row.get("Metric Name", "gpu__time").startswith("gpu__time"):
            continue  # (captures with extra metrics: only the durations are summed here)
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("itb::", "")
        agg[name][0] += 1
        agg[name][1] += v
        by_grid[(name.split("<")[0], row.get("Grid Size", ""))][0] += 1
        by_grid[(name.split("<")[0], row.get("Grid Size", ""))][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are "
                "cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {v[1] / tot * 100:.1f}% |\n")
        f.write("\n| kernel | grid | launches | avg us |\n|---|---|---|---|\n")
        for (k, g), v in sorted(by_grid.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {g} | {v[0]} | {v[1] / v[0]:.2f} |\n")


WANT = ["Kernel Name", "Grid Size", "Block Size", "launch__cluster_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows[2:]:
            f.write("```\n")
            for w in WANT:
                if w in idx:
                    f.write(f"{w:75s} {r[idx[w]]} {units[idx[w]]}\n")
            f.write("```\n\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
