#!/usr/bin/env python
"""Summarise an .ncu-rep (captured with --set full --import-source on): headline metrics per kernel, then for one kernel
the stall mix and the source lines that collect the most stall samples / executed instructions (ncu's own CUDA-source
correlation: no SASS offsets to match against a local build).
usage: python profiles/analyze_ncu.py report.ncu-rep [kernel-substring] [top-n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kname = sys.argv[2] if len(sys.argv) > 2 else "k_lookup"
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 22
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "dram__sectors_read.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__occupancy_limit_registers"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:60])
    for w in want:
        if w in hdr:
            print(f"  {w:72s} {r[hdr.index(w)]} {units[hdr.index(w)]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, hdr, lines, seen = None, None, [], False
for r in rows:
    if not r:
        continue
    if r[0] == "Function Name":
        if seen and lines:
            break
        cur = r[1]
        seen = kname in cur
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if seen and hdr and r[0].isdigit():
        lines.append(r)
if not lines:
    sys.exit(f"no source rows for a kernel matching {kname!r}")
H = {}
for i, h in enumerate(hdr):
    H.setdefault(h, i)
samp, inst = H["# Samples"], H["Instructions Executed"]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0



stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(num(r[H[s]]) for r in lines) for s in stalls}
T = sum(tot.values()) or 1
print(f"\n== {cur}")
print("stall mix:", ", ".join(f"{s[6:]} {100 * v / T:.1f}%" for s, v in sorted(tot.items(), key=lambda x: -x[1])[:8]))
ts, ti = sum(num(r[samp]) for r in lines) or 1, sum(num(r[inst]) for r in lines) or 1
print(f"{'line':>5} {'samples%':>8} {'inst%':>6}  top stall      source")
for r in sorted(lines, key=lambda r: -num(r[samp]))[:topn]:
    st = max(stalls, key=lambda s: num(r[H[s]]))
    print(f"{r[0]:>5} {100 * num(r[samp]) / ts:8.1f} {100 * num(r[inst]) / ti:6.1f}  {st[6:]:<14} {r[1].strip()[:110]}")
