"""SASS opcode census of the in-tree library (cuobjdump -sass): the tensor / TMA / TMEM / mbarrier instructions that prove
the tcgen05 path, in total and per kernel.  usage: python tools/sass_census.py > profiles/sass_census_r2.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "keisei_b200" / "libkeisei_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
pat = re.compile(r"\b(UTCHMMA[.\w]*|UTCQMMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMACCTL[.\w]*|UBLKCP[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|"
                 r"SYNCS[.\w]*|ELECT|STAS[.\w]*|HMMA[.\w]*|UCGABAR[.\w]*)")
total = collections.Counter()
per_kernel = collections.Counter()
kernel = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kernel = m.group(1)
        continue
    m = pat.search(line)
    if m and "/*" in line:
        op = re.sub(r"\.(U?[0-9]+x[0-9]+b?|x[0-9]+|[0-9]+dp[0-9]+bit\w*)$", ".", m.group(1))
        op = "LDTM." if op.startswith("LDTM") else op
        total[op] += 1
        if re.match(r"UTCHMMA|UTMALDG|UBLKCP|UTCBAR|LDTM|STAS", op):
            per_kernel[kernel] += 1
print(f"# SASS opcode census of keisei_b200/libkeisei_b200.so (cuobjdump -sass, sm_100a), round 2")
print(f"# built from HEAD {head}; tcgen05.mma -> UTCHMMA, cta_group::2 -> .2CTA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP, tcgen05.commit -> UTCBAR, st.async -> STAS")
for op, n in total.most_common():
    print(f"{n:7d} {op}")
print("\n# per kernel (UTCHMMA + UTMALDG + UBLKCP + UTCBAR + LDTM + STAS instructions)")
demangle = subprocess.run(["c++filt"], input="\n".join(per_kernel), capture_output=True, text=True).stdout.splitlines()
for (k, n), d in sorted(zip(per_kernel.items(), demangle), key=lambda t: -t[0][1]):
    print(f"{n} {d}")
