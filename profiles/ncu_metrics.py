"""Selected raw metrics of every kernel in an .ncu-rep as JSON lines:
    python profiles/ncu_metrics.py report.ncu-rep substring [substring ...]"""
import csv, io, json, subprocess, sys
rep, wants = sys.argv[1], sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    out = {"kernel": r[hdr.index("Kernel Name")]}
    for h, u, v in zip(hdr, units, r):
        if any(w in h for w in wants):
            try:
                out[h + (" [%s]" % u if u else "")] = float(v.replace(",", ""))
            except ValueError:
                out[h] = v
    print(json.dumps(out))
