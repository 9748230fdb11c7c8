"""Per-kernel SASS opcode histogram of libtpx.so (the mnemonics that prove tcgen05 / TMEM / bulk-copy use:
UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk).
usage: python scripts/sass_summary.py [lib] > profiles/rNN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "torchphysics_b200/lib/libtpx.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKPF", "SYNCS", "USETMAXREG", "REDUX", "RED", "ATOMS", "SHFL", "MUFU")
print(f"# SASS opcode summary of {lib} (cuobjdump -sass); tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM, tcgen05.commit = UTCBAR,")
print("# cp.async.bulk = UBLKCP, cp.async.bulk.prefetch = UBLKPF, mbarrier = SYNCS, setmaxnreg = USETMAXREG")
tot = collections.Counter()
for k in sorted(hist):
    h = hist[k]
    n = sum(h.values())
    if n == 0:
        continue
    short = re.sub(r"\(.*", "", k)
    keys = " ".join(f"{op}={h[op]}" for op in KEY if h[op])
    top = " ".join(f"{op}:{c}" for op, c in h.most_common(6))
    print(f"{short}  [{n} instr]  {keys}  | top: {top}")
    tot.update(h)
print("TOTAL " + " ".join(f"{op}={tot[op]}" for op in KEY))
