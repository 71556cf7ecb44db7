"""Per-stage stall breakdown of a fused-kernel ncu capture: splits the SASS at RET/EXIT (one segment per generated stage
function), prints the stall-reason totals and the hottest instructions of every segment.
usage: python scripts/ncu_segments.py prof.ncu-rep [top_n]"""
import csv, io, subprocess, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 4
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]; data = rows[2:]
isrc = hdr.index("Source"); isamp = hdr.index("Warp Stall Sampling (All Samples)"); iexe = hdr.index("Instructions Executed")
stallcols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
fid = 0; segs = {}
for r in data:
    segs.setdefault(fid, []).append(r)
    if r[isrc].strip().startswith("RET") or "EXIT" in r[isrc]: fid += 1
total = sum(int(r[isamp] or 0) for r in data); texe = sum(int(r[iexe] or 0) for r in data)
print(f"{path}: {total} samples, {texe} warp instructions")
for k, seg in segs.items():
    samp = sum(int(r[isamp] or 0) for r in seg); exe = sum(int(r[iexe] or 0) for r in seg)
    if samp < total * 0.01: continue
    tot = {}
    for r in seg:
        for i in stallcols:
            try: tot[hdr[i][6:]] = tot.get(hdr[i][6:], 0) + int(r[i] or 0)
            except ValueError: pass
    reasons = ", ".join(f"{n} {v}" for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:5])
    ops = {}
    for r in seg:
        op = r[isrc].split()[0 if not r[isrc].strip().startswith("@") else 1].split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    mix = " ".join(f"{o}:{n}" for o, n in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
    print(f"seg {k:2d}: {len(seg):4d} instr  samples {samp:6d} ({100*samp/total:4.1f}%)  executed {100*exe/texe:4.1f}%  [{reasons}]  {mix}")
    for r in sorted(seg, key=lambda r: -int(r[isamp] or 0))[:topn]:
        print(f"        {r[isamp]:>6} {r[isrc][:80]}")
