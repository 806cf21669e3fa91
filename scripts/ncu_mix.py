"""Summarise an `ncu --page source --csv` export: executed warp-instructions by opcode, stall reasons."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
si, ei, smp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
ops, samples, stalls, tot = collections.Counter(), collections.Counter(), collections.Counter(), 0
for r in rows[hdr + 1:]:
    if len(r) <= ei or not r[ei].isdigit():
        continue
    src = r[si].strip().split()
    if not src:
        continue
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    n = int(r[ei])
    ops[op] += n
    tot += n
    samples[op] += int(r[smp])
    for i in stall_cols:
        stalls[h[i]] += int(r[i])
print("total warp instructions", tot)
for k, v in ops.most_common(24):
    print(f"{k:8s} {v:10d} {100 * v / tot:5.1f}%  samples {samples[k]}")
ts = sum(stalls.values())
for k, v in stalls.most_common(10):
    print(f"{k:28s} {100 * v / ts:5.1f}%")
