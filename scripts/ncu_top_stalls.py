"""Print the SASS instructions with the most warp-stall samples from `ncu --page source --csv` output."""
import csv
import sys

path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
print(rows[0][1][:150])
data = rows[2:]
tot = sum(int(r[2]) for r in data if r[2].isdigit())
print("total samples", tot)
top = sorted([(int(r[2]), i, r[1].strip()) for i, r in enumerate(data) if r[2].isdigit()], reverse=True)[:n]
for s, i, src in sorted(top, key=lambda t: t[1]):
    print(f"{i:5d} {s:6d} {100 * s / tot:5.1f}%  {src[:120]}")
