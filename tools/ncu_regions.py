#!/usr/bin/env python
"""Development helper: split an `ncu --page source --csv --print-source sass` export of surface_warp_kernel
into the wave loop (between the first MUFU of the loop and its closing BRA.U) and everything else, and
print samples / instructions per region plus the top stall reasons."""
import csv, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path)))
# first kernel only
start = 1
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) < len(hdr) // 2:
        break
    data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
A, S, N, IE = ci["Address"], ci["Source"], ci["# Samples"], ci["Instructions Executed"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][A], 16)
# find loop regions: backward branches
regions = []
for k, r in enumerate(data):
    s = r[S].strip()
    if s.startswith("BRA") or " BRA" in s:
        tgt = s.split()[-1].rstrip(";")
        try:
            t = int(tgt, 16)
        except ValueError:
            continue
        if t >= base:
            t -= base
        off = int(r[A], 16) - base
        if t < off:
            regions.append((t, off, int(r[IE] or 0)))
tot_s = sum(int(r[N] or 0) for r in data); tot_i = sum(int(r[IE] or 0) for r in data)
print("total samples", tot_s, "total warp-instructions", tot_i)
for t, off, cnt in regions:
    sel = [r for r in data if t <= int(r[A], 16) - base <= off]
    s = sum(int(r[N] or 0) for r in sel); i = sum(int(r[IE] or 0) for r in sel)
    st = collections.Counter()
    for r in sel:
        for c in stall_cols:
            st[c] += int(r[ci[c]] or 0)
    print(f"loop [{t:#x},{off:#x}] n_instr={len(sel)} trips={cnt} samples={s} ({100*s/tot_s:.1f}%) instr={i} ({100*i/tot_i:.1f}%)", st.most_common(6))
if len(sys.argv) > 2:
    lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
    for r in data:
        o = int(r[A], 16) - base
        if lo <= o <= hi:
            top = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
            print(f"{o:#06x} {r[N]:>6} {r[IE]:>9}  {r[S].strip():70s} {[(c, n) for n, c in top if n]}")
