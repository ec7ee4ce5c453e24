"""profiles/ helper: tensor-core / TMEM / bulk-copy / mbarrier mnemonics per kernel from `cuobjdump -sass` of the built
library (the SASS evidence that the hot kernels are tcgen05 / TMA code).  usage: python tools/sass_extract.py > profiles/sass_tcgen05_r2.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gapartnet_b200", "libgapart_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCBAR|UTCATOMSWS[.\w]*|STTM[.\w]*|LDTM[.\w]*|UBLKCP[.\w]*|SYNCS[.\w]*|LDGSTS[.\w]*|LDGSTSBAR[.\w]*|REDUX|UTMALDG[.\w]*)\b")
per = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        per[name] = collections.Counter()
        continue
    if name:
        m = pat.search(line)
        if m:
            per[name][m.group(1)] += 1
print("# cuobjdump -sass gapartnet_b200/libgapart_b200.so: tensor-core / TMEM / bulk-copy / mbarrier mnemonics per kernel (round 2)")
print("# UTCHMMA = tcgen05.mma, STTM/LDTM = tcgen05.st/ld, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops")
for k, c in per.items():
    if c.get("UTCHMMA") or any(x.startswith("UBLKCP") for x in c):
        print(f"{k}: " + ", ".join(f"{m} x{n}" for m, n in sorted(c.items())))
