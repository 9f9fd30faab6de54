#!/usr/bin/env python
"""Summarise an .ncu-rep of k_classify: headline metrics, stall mix, instructions / stall samples per source line.
usage: python profiles/analyze_ncu.py gpurun_out/prof.ncu-rep [libkuq.so]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

rep = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "krakenuniq_b200", "lib", "libkuq.so")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "dram__sectors_read.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__occupancy_limit_registers"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:60])
    for w in want:
        if w in hdr:
            print(f"  {w:72s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
H = {h: i for i, h in enumerate(hdr)}
kname = sys.argv[3] if len(sys.argv) > 3 else "k_lookupILi0"
short = re.sub(r"ILi\d+.*", "", kname)
data, active = [], False
for r in rows:
    if r and r[0] == "Kernel Name":
        if data:
            break
        active = short in r[1]
        continue
    if active and len(r) > 10 and r[0].startswith("0x"):
        data.append(r)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(int(r[H[s]] or 0) for r in data) for s in stalls}
T = sum(tot.values()) or 1
print("stall mix:", ", ".join(f"{s[6:]} {100 * v / T:.1f}%" for s, v in sorted(tot.items(), key=lambda x: -x[1])[:7]))
# map SASS offsets to source lines
tmp = "/tmp/kuq_cub"
os.makedirs(tmp, exist_ok=True)
subprocess.run(f"cd {tmp} && rm -f *.cubin && cuobjdump -xelf all {so} >/dev/null 2>&1", shell=True)
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, "kuq_kernels.sm_100a.cubin")], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if ".text." in l and kname in l][0]
cur, m = None, {}
for l in sass[start + 1:]:
    if l.startswith("//---") and m:
        break
    a = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if a:
        cur = (a.group(1).split("/")[-1], int(a.group(2)))
        continue
    b = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*);", l)
    if b:
        m[int(b.group(1), 16)] = cur
base = int(data[0][0], 16)
inst, samp = collections.Counter(), collections.Counter()
for r in data:
    k = m.get(int(r[0], 16) - base)
    inst[k] += int(r[H["Instructions Executed"]])
    samp[k] += int(r[H["# Samples"]])
ti, ts = sum(inst.values()), sum(samp.values())
srcl = open(os.path.join(os.path.dirname(so), "..", "csrc", "kuq_kernels.cu")).read().split("\n")
print(f"instructions executed {ti}; samples {ts}")
print("top lines by stall samples:")
for k, v in samp.most_common(22):
    txt = srcl[k[1] - 1].strip()[:74] if k and k[0] == "kuq_kernels.cu" else str(k)
    print(f"  {100 * v / ts:5.1f}% samp {100 * inst[k] / ti:5.1f}% inst  L{k[1] if k else 0}: {txt}")
