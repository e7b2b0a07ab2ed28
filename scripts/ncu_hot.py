#!/usr/bin/env python
"""Read a .ncu-rep here (no GPU): headline metrics, stall mix, hottest SASS lines, instruction count per executed-count class.
    python scripts/ncu_hot.py gpurun_out/prof_x.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 22


def page(name, *extra):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def f(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


raw = page("raw")
hdr, r = raw[0], raw[2]
for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread"]:
    if k in hdr:
        print(f"{k:70s} {r[hdr.index(k)]}")
st = [(hdr[i][len("smsp__pcsamp_warps_issue_stalled_"):], f(r[i])) for i in range(len(hdr))
      if "pcsamp_warps_issue_stalled" in hdr[i] and "not_issued" not in hdr[i]]
st.sort(key=lambda t: -t[1])
print("stalls:", ", ".join(f"{k}={int(v)}" for k, v in st[:9]))

src = page("source", "--print-source", "sass")
h = src[1]
i_src, i_s, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
cols = {k: h.index(k) for k in h if k.startswith("stall_") and "Not Issued" not in k}
body = []
for row in src[2:]:
    if len(row) < len(h) - 5 or row[0] in ("Kernel Name", "Address"):
        break
    body.append(row)
tot = sum(f(b[i_s]) for b in body)
print(f"lines {len(body)}  samples {int(tot)}  warp-instr {int(sum(f(b[i_ex]) for b in body))}")
cls = collections.Counter()
for b in body:
    cls[int(f(b[i_ex]))] += 1
print("executed-count classes (count -> #lines):", sorted(((k, v) for k, v in cls.items() if v >= 12), key=lambda t: -t[1])[:8])
top = sorted(range(len(body)), key=lambda i: -f(body[i][i_s]))[:topn]
for i in sorted(top):
    b = body[i]
    why = " ".join(f"{k[6:]}={b[c]}" for k, c in cols.items() if f(b[c]) > 0.15 * max(f(b[i_s]), 1))
    print(f"{i:5d} {b[i_src][:64]:64s} smp {b[i_s]:>6s} ex {b[i_ex]:>9s} {why}")
