"""Summarise an .ncu-rep: key raw metrics + top stalled SASS instructions (run where ncu is installed).
    python profiles/ncu_summ.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_uniform", "launch__grid_size", "launch__block_size"]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for h, u, v in zip(hdr, units, r):
        if any(h == k or h.startswith(k + ".") and k.endswith("sum") and h.endswith("per_second") for k in KEYS) or "stalled" in h and "per_issue_active" in h and float(v or 0) > 0.3:
            print("  %-95s %-12s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"] + [len(rows)]
for sec in range(len(heads) - 1):          # one section per profiled kernel
    hdr = rows[heads[sec]]
    data = [r for r in rows[heads[sec] + 1:heads[sec + 1]] if len(r) == len(hdr) and r[0] != "Address"]
    iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[iS]) for r in data)
    print("-- source section %d: total samples %d, instructions %d" % (sec, tot, len(data)))
    for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:top]):
        r = data[i]
        st = sorted(((int(r[c]), hdr[c]) for c in stall), reverse=True)[:2]
        print("%5d %6.2f%% exec=%-8s %-60s %s" % (i, 100.0 * int(r[iS]) / max(tot, 1), r[iEx], r[iSrc].strip()[:60], st))
