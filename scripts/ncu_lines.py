"""Development aid: per-source-line instruction counts and stall samples of an ncu report (needs -lineinfo).
usage: python scripts/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None
tot_i = tot_s = 0
lines = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Name", "File Path"):
        cur = r[1].split("/")[-1]
        hdr = None
        continue
    if r and r[0] == "Line No":
        hdr = {}
        for i, h in enumerate(r):
            hdr.setdefault(h, i)
        continue
    if cur is None or hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    try:
        n = int(r[hdr["Instructions Executed"]]); s = int(r[hdr["# Samples"]])
    except (KeyError, ValueError):
        continue
    tot_i += n; tot_s += s
    lines.append((n, s, cur, int(r[0]), r[1].strip()))
print(f"total {tot_i} warp instructions, {tot_s} samples")
for n, s, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100.0 * n / tot_i:5.1f}% inst {100.0 * s / max(tot_s, 1):5.1f}% samp  {f}:{ln}  {src[:100]}")
