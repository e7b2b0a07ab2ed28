#!/usr/bin/env python
"""Turn what a GPU visit left in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarise_profiles.py r1b           # tag = round / visit label

  launches.csv (ncu --metrics gpu__time_duration.sum of `bench.py --steps 1 --warmup 3 --no-cpu-baseline`)
      -> profiles/<tag>_launches_step.csv   per-kernel launches / total us / share of ONE timed step
  prof_*.ncu-rep (ncu --set full)           -> profiles/<tag>_<name>_ncu.csv   selected raw metrics per captured launch
"""
import collections
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_selected"]


def launches(tag, path=None, name="launches_step"):
    path = path or os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if n.startswith("graph_ptr_kernel")]
    # bench.py --steps 1 --warmup 3: steps 0-2 warm-up, step 3 = the timed one
    lo, hi = (starts[3], starts[4]) if len(starts) > 4 else (0, len(rows))
    agg = collections.OrderedDict()
    for r in rows[lo:hi]:
        n = r["Kernel Name"].split("(")[0][:90]
        t = float(r["Metric Value"].replace(",", ""))
        t = t / 1e3 if r["Metric Unit"].startswith("n") else t
        c = agg.setdefault(n, [0, 0.0, r["Grid Size"], r["Block Size"]])
        c[0] += 1
        c[1] += t
    tot = sum(v[1] for v in agg.values())
    dst = os.path.join(PROF, f"{tag}_{name}.csv")
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 3 "
                    "--no-cpu-baseline ; the one timed step (cold-cache, serialised: compare shares)"])
        w.writerow(["kernel", "launches", "total_us", "share", "avg_us", "grid", "block"])
        for n, (c, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([n, c, f"{t:.1f}", f"{t / tot:.4f}", f"{t / c:.1f}", g, b])
        w.writerow(["TOTAL", hi - lo, f"{tot:.1f}", "1.0", "", "", ""])
    print("wrote", dst, f"({hi - lo} launches, {tot / 1e3:.2f} ms)")


def full(tag):
    for rep in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        cols = [hdr.index(k) for k in KEEP if k in hdr]
        dst = os.path.join(PROF, f"{tag}_{name}_ncu.csv")
        with open(dst, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow([f"# ncu --set full --clock-control none --import-source on ({os.path.basename(rep)}); "
                        "one column per captured launch"])
            w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
            for c in cols:
                w.writerow([hdr[c], units[c]] + [r[c] for r in rows[2:]])
        print("wrote", dst)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    full(tag)
