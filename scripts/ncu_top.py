#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + hottest SASS instructions with their stall reasons."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:70s} {vals[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
S = lambda r: int(r[ci["# Samples"]]); E = lambda r: int(r[ci["Instructions Executed"]])
tot = sum(S(r) for r in data)
print("samples", tot, "instructions", len(data), "warp-instr executed", sum(E(r) for r in data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -S(r))[:topn]:
    st = sorted(((int(r[ci[h]]), h.replace("stall_", "")) for h in stalls), reverse=True)[:2]
    print(f"{S(r):6d} {S(r)/tot*100:5.1f}% idx={data.index(r):5d} exec={E(r):>9d} {r[ci['Source']].strip()[:64]:64s} {[x for x in st if x[0]]}")
