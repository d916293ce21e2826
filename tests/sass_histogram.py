"""SASS evidence: instruction histogram per kernel of libpe_b200.so (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = bulk async
copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier).  Usage: python tests/sass_histogram.py > profiles/r2_sass_histogram.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "playableenvironments_b200", "libpe_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
hist, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        hist[cur][m.group(1).split(".")[0]] += 1
names = subprocess.run(["cu++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS", "RED", "ATOMG", "SHFL"]
print("# SASS instruction histogram per kernel (cuobjdump -sass playableenvironments_b200/libpe_b200.so)\n")
print("| kernel | instructions | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
for (k, c), name in sorted(zip(hist.items(), names), key=lambda kv: -kv[0][1].get("UTCHMMA", 0)):
    name = re.sub(r"\((Pe|const|float|int|unsigned|long|double|\)).*", "", name).replace("(anonymous namespace)::", "").replace("void ", "")
    print("| `" + name[:80] + "` | " + str(sum(c.values())) + " | " + " | ".join(str(c.get(x, 0)) for x in keys) + " |")
