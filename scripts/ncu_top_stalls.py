"""Top SASS instructions by warp-stall samples from `ncu -i rep --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, src, ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
body = rows[2:]
tot = sum(int(r[si] or 0) for r in body)
top = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
print("total samples", tot)
for i in sorted(top):
    r = body[i]
    print(f"{i:5d} {int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}%  exec={r[ex]:>9s} {r[src].strip()[:110]}")
