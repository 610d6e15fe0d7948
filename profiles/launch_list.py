"""Print the last training step of an ncu launch list (gpu__time_duration.sum, --csv)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
out = []
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
    out.append((r[ki].split("(")[0], v))
idx = [i for i, (n, _) in enumerate(out) if "pack_weights" in n]
last = out[idx[-1]:]
for n, v in last:
    print("%-40s %8.1f us" % (n, v))
print("%-40s %8.1f us" % ("sum", sum(v for _, v in last)))
