#!/usr/bin/env python
"""Per-source-line stall samples of an .ncu-rep (needs -lineinfo + --import-source on): where the warps wait."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr, ci = None, None, None
agg = collections.defaultdict(lambda: [0, 0, "", collections.Counter()])
tot = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ci = {h: i for i, h in enumerate(hdr)}; si = hdr.index("# Samples"); ei = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] != "-":      # SASS row under a source line
        continue
    try: s = int(r[si]); e = int(r[ei])
    except ValueError: continue
    k = (fname, int(r[0]))
    a = agg[k]; a[0] += s; a[1] += e; a[2] = r[1].strip()
    for h, i in ci.items():
        if h.startswith("stall_") and "Not Issued" not in h:
            try: a[3][h[6:]] += int(r[i])
            except ValueError: pass
    tot += s
print("total samples", tot)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{a[0]:7d} {100*a[0]/max(tot,1):5.1f}% exec={a[1]:>10d} {k[0]}:{k[1]:<4d} {a[2][:70]:70s} {a[3].most_common(2)}")
