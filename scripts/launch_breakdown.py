"""Prints the kernels of the LAST batch in an ncu gpu__time_duration launch list (csv)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
out = [(r[ki][:56], float(r[vi].replace(",", ""))) for r in rows[1:]]
start = max((i for i, (n, _) in enumerate(out) if sys.argv[2] in n), default=0) if len(sys.argv) > 2 else 0
for n, v in out[start:]:
    print(f"{n:58s} {v / 1000:9.1f} us")
print(f"total {sum(v for _, v in out[start:]) / 1e6:.3f} ms")
