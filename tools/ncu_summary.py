#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline metrics, opcode mix, stall reasons (whole kernel and
for a given address window).  usage: ncu_summary.py rep [lo_suffix hi_suffix]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_pipe_tmem.sum", "launch__grid_size", "launch__block_size"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print("%-75s %-10s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isrc, iex, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows[2:] if len(r) > iex]
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def report(sel, title):
    tot = sum(int(r[iex] or 0) for r in sel); ts = sum(int(r[isamp] or 0) for r in sel)
    print("== %s: %d SASS instrs, %d executed, %d samples" % (title, len(sel), tot, ts))
    c = Counter(); 
    for r in sel:
        s = r[isrc]; op = s.split()[1] if s.startswith("@") else s.split()[0]
        c[op] += int(r[iex] or 0)
    print("   opcodes:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tot, 1)) for k, v in c.most_common(14)))
    st = {h: sum(int(r[hdr.index(h)] or 0) for r in sel) for h in stall}
    print("   stalls :", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(ts, 1)) for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]))
report(data, "whole kernel")
# automatic windows: group instructions by identical execution count (one bucket per loop nest / role)
from collections import defaultdict
buckets = defaultdict(list)
for r in data:
    buckets[int(r[iex] or 0)].append(r)
for cnt, sel in sorted(buckets.items(), key=lambda kv: -kv[0] * len(kv[1]))[:5]:
    report(sel, "instructions executed %d times" % cnt)
if len(sys.argv) > 3 and not sys.argv[2].startswith("--"):
    lo = [int(r[ia], 16) for r in data if r[ia].endswith(sys.argv[2])][0]
    hi = [int(r[ia], 16) for r in data if r[ia].endswith(sys.argv[3])][0]
    report([r for r in data if lo <= int(r[ia], 16) <= hi], "window %s..%s" % (sys.argv[2], sys.argv[3]))
if "--list" in sys.argv:
    for r in data:
        if int(r[isamp] or 0) >= int(sys.argv[sys.argv.index("--list") + 1]):
            st = {h: int(r[hdr.index(h)] or 0) for h in stall}
            print(r[ia][-5:], r[isamp].rjust(5), r[iex].rjust(9), r[isrc][:64].ljust(64), sorted(st.items(), key=lambda x: -x[1])[:2])
