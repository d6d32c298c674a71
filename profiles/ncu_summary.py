import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = [("Kernel Name", 34), ("gpu__time_duration.sum", 9), ("dram__bytes_read.sum", 9), ("dram__bytes_write.sum", 9),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 7), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 7),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", 7), ("launch__registers_per_thread", 5), ("smsp__inst_executed.sum", 12),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", 7), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 10),
        ("lts__t_sector_hit_rate.pct", 7)]
idx = [(hdr.index(w) if w in hdr else -1, n) for w, n in want]
print(" | ".join(w.split(".")[0][-n:].ljust(n) for (w, n) in want))
print(" | ".join((rows[1][i][:n] if i >= 0 else "NA").ljust(n) for i, n in idx))
for r in rows[2:]:
    print(" | ".join((r[i][:n] if i >= 0 else "NA").ljust(n) for i, n in idx))
