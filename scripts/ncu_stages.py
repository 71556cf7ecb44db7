"""Attribute warp-stall samples of a fused-kernel ncu capture to the generated stage functions (split at RET/EXIT).
usage: python scripts/ncu_stages.py prof.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
isrc, isamp = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[isamp] or 0) for r in data)
acc, fid = {}, 0
for n, r in enumerate(data):
    s = int(r[isamp] or 0)
    a = acc.setdefault(fid, [0, 0, n, []])
    a[0] += s; a[1] += 1; a[3].append((s, r[isrc][:64]))
    if r[isrc].strip().startswith("RET") or "EXIT" in r[isrc]: fid += 1
print("total samples", tot)
for f, (s, n, start, ins) in acc.items():
    if s < 0.01 * tot: continue
    nd = sum(1 for x in ins if "DFMA" in x[1] or "DMUL" in x[1]); ng = sum(1 for x in ins if "LDG" in x[1] or x[1].strip().startswith("LD.E"))
    print(f"segment {f:3d} start {start:5d} ninstr {n:5d} samples {s:7d} ({100*s/tot:4.1f}%)  DFMA {nd} LDG {ng}")
    for t in sorted(ins, reverse=True)[:3]: print("      ", t)
