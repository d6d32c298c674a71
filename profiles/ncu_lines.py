"""Per-source-line samples / instructions of one kernel: python profiles/ncu_lines.py rep kernel [top]"""
import csv, subprocess, sys
rep, name = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", name, "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
cs, ci = hdr.index("# Samples"), hdr.index("Instructions Executed")
sb = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
lines = []
for r in rows[hi + 1:]:
    if len(r) > ci and r[0].isdigit():
        st = sorted(((int(r[i] or 0), hdr[i]) for i in sb), reverse=True)[:3]
        lines.append((int(r[cs] or 0), int(r[ci] or 0), r[0], r[1].strip()[:90], ", ".join(f"{h[6:]}={v}" for v, h in st if v)))
    if r and r[0] == "File Path" and lines:
        break
tot = sum(l[0] for l in lines) or 1
toti = sum(l[1] for l in lines) or 1
print(f"total samples {tot}, warp instructions {toti}")
for s, i, ln, src, st in sorted(lines, reverse=True)[:top]:
    print(f"{100*s/tot:5.1f}% smp {100*i/toti:5.1f}% inst  L{ln:>4s}  {src:90s} {st}")
