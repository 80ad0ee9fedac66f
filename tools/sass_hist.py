"""Opcode histogram (weighted by executed count) from `ncu -i rep --page source --csv --kernel-name regex:K` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r:
        hdr, start = r, i + 1
        break
idx = {h: i for i, h in enumerate(hdr)}
src = idx["Source"]
ex = [i for h, i in idx.items() if "Instructions Executed" in h and "Thread" not in h][0]
ops, total = collections.Counter(), 0
for r in rows[start:]:
    try:
        n = int(r[ex])
    except Exception:
        continue
    toks = r[src].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    ops[op.split(".")[0]] += n
    total += n
print("total warp-instructions", total)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print("%-12s %14d %5.1f%%" % (op, n, 100.0 * n / total))
