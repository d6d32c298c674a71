"""Per-source-line samples / instructions of one kernel: python profiles/ncu_lines.py rep kernel [top]"""
import csv, os, subprocess, sys
rep, name = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", name, "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines, hdr, fname, seen = [], None, "", set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
        continue
    if r[0] == "Function Name":
        if len(seen) and (fname, "hdr") in seen:
            pass
        continue
    if r[0] == "Line No":
        hdr = r
        cs, ci = hdr.index("# Samples"), hdr.index("Instructions Executed")
        sb = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        key = (fname, r[0])
        if key in seen:  # a second launch of the same kernel: keep the first
            continue
        seen.add(key)
        try:
            int(r[cs] or 0), int(r[ci] or 0)
        except ValueError:
            continue
        st = sorted(((int(r[i] or 0), hdr[i]) for i in sb), reverse=True)[:3]
        lines.append((int(r[cs] or 0), int(r[ci] or 0), f"{fname[:12]}:{r[0]}", r[1].strip()[:86], ", ".join(f"{h[6:]}={v}" for v, h in st if v)))
tot = sum(l[0] for l in lines) or 1
toti = sum(l[1] for l in lines) or 1
print(f"total samples {tot}, warp instructions {toti}")
for s, i, ln, src, st in sorted(lines, reverse=True)[:top]:
    print(f"{100*s/tot:5.1f}% smp {100*i/toti:5.1f}% inst  {ln:>17s}  {src:86s} {st}")
