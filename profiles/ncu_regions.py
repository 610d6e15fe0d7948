"""Executed warp-instructions and stall samples of one profiled kernel, summed per callee region (split at RET/EXIT):
    python profiles/ncu_regions.py report.ncu-rep [section]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"] + [len(rows)]
hdr = rows[heads[want]]
data = [r for r in rows[heads[want] + 1:heads[want + 1]] if len(r) == len(hdr) and r[0] != "Address"]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
regions, cur = [], [0, 0, 0, 0]
for i, r in enumerate(data):
    cur[1] += 1; cur[2] += int(r[iEx]); cur[3] += int(r[iS])
    if r[iSrc].strip().startswith(("RET", "EXIT")) and cur[1] > 40:
        regions.append(cur + [i]); cur = [i + 1, 0, 0, 0]
tot_e, tot_s = sum(x[2] for x in regions) or 1, sum(x[3] for x in regions) or 1
for st, n, e, s_, end in regions:
    print("instr %5d..%5d  static %5d  executed %10d (%4.1f%%)  samples %6d (%4.1f%%)" % (st, end, n, e, 100.0 * e / tot_e, s_, 100.0 * s_ / tot_s))
# hottest 64-instruction windows by executed count
win = 64
tops = sorted(((sum(int(r[iEx]) for r in data[i:i + win]), i) for i in range(0, len(data), win)), reverse=True)[:12]
for e, i in sorted(tops, key=lambda x: x[1]):
    print("window %5d: executed %9d  e.g. %s" % (i, e, data[i][iSrc].strip()[:50]))
