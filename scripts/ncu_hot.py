"""Development aid: hottest SASS instructions of an ncu report (source page), with opcode histogram.
usage: python scripts/ncu_hot.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot_s = sum(int(r[ci["# Samples"]]) for r in body)
tot_i = sum(int(r[ci["Instructions Executed"]]) for r in body)
print(f"{len(body)} SASS instructions, {tot_s} samples, {tot_i} warp instructions executed")
ops = collections.Counter()
opsamp = collections.Counter()
for r in body:
    op = r[ci["Source"]].split()[0] if not r[ci["Source"]].strip().startswith("@") else r[ci["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] += int(r[ci["Instructions Executed"]])
    opsamp[op] += int(r[ci["# Samples"]])
print("opcode: executed share | sample share")
for op, n in ops.most_common(28):
    print(f"  {op:10s} {100.0 * n / tot_i:5.1f} %  {100.0 * opsamp[op] / max(tot_s, 1):5.1f} %")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("hottest instructions:")
for r in sorted(body, key=lambda r: -int(r[ci["# Samples"]]))[:top]:
    st = sorted(((int(r[ci[h]]), h) for h in stall_cols), reverse=True)[:2]
    print(f"  {int(r[ci['# Samples']]):6d} {r[ci['Address']][-5:]} {r[ci['Source']].strip()[:70]:70s} {st}")
