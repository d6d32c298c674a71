"""Dump the metrics of one kernel (substring match) from an .ncu-rep: python profiles/ncu_kernel.py rep name [filter]"""
import csv, subprocess, sys
rep, name = sys.argv[1], sys.argv[2]
flt = sys.argv[3:] or ["stall", "duration", "dram__bytes", "throughput", "occupancy", "warps_active", "registers", "inst_executed.sum", "bank_conflict", "hit_rate", "sectors_per_request", "lsu", "issue"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    if name in r[ki]:
        print("==", r[ki][:80])
        for h, u, v in zip(hdr, units, r):
            if any(f in h for f in flt) and v not in ("", "0", "n/a"):
                print(f"  {h:90s} {v:>16s} {u}")
        break
