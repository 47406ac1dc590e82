"""Aggregate an `ncu --page source --csv` dump (SASS view): stall samples and executed instructions per opcode,
plus the hottest instructions.   python tools/ncu_source_top.py file.csv[.gz] [top_n]"""
import collections
import csv
import gzip
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
names = rows[hdr]
ix = {n: i for i, n in enumerate(names)}
ops = collections.defaultdict(lambda: [0, 0])
lines = []
tot_s = tot_i = 0
for k, r in enumerate(rows[hdr + 1:]):
    if len(r) < len(names):
        continue
    s = int(r[ix["# Samples"]] or 0)
    n = int(r[ix["Instructions Executed"]] or 0)
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.rstrip(";")
    ops[op][0] += s
    ops[op][1] += n
    tot_s += s
    tot_i += n
    lines.append((s, n, k, src))
print(f"total samples {tot_s}, warp instructions {tot_i}")
print("-- by opcode (samples %, instructions %)")
for op, (s, n) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {op:28s} {100 * s / max(tot_s, 1):6.2f} %  {100 * n / max(tot_i, 1):6.2f} %  ({n})")
print("-- hottest instructions")
for s, n, k, src in sorted(lines, reverse=True)[:top]:
    print(f"  #{k:5d} {100 * s / max(tot_s, 1):6.2f} %  x{n:<12d} {src}")
