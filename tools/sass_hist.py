"""Aggregates an `ncu --page source --csv` dump (SASS view): executed warp-instructions and stall samples by opcode,
and the hottest instructions.  usage: sass_hist.py file.csv [top_n]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = next(r for r in rows if r and r[0] == "Address")
idx = {h: i for i, h in enumerate(hdr)}
tot = 0; byop = collections.Counter(); samp = collections.Counter(); lines = []
for r in rows:
    if len(r) < 10 or r[0] == "Address":
        continue
    ins = r[idx["Source"]].split()
    op = ins[0] if not ins[0].startswith("@") else ins[1]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "ATOM", "LDGSTS")) else op.split(".")[0]
    n = int(r[idx["Instructions Executed"]]); s = int(r[idx["# Samples"]])
    byop[op] += n; samp[op] += s; tot += n
    lines.append((s, n, r[idx["Source"]].strip(), r[idx["L1 Wavefronts Shared"]] if "L1 Wavefronts Shared" in idx else ""))
print("total warp-instr", tot, "samples", sum(samp.values()))
for op, n in byop.most_common(topn):
    print(f"{op:14s} {n:12d} {100*n/tot:5.1f}%  samples {samp[op]}")
print("-- hottest by samples")
for s, n, src, wf in sorted(lines, reverse=True)[:topn]:
    print(s, n, wf, src)
