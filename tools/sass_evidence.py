"""Regenerates profiles/r2_sass_tcgen05.txt: per-kernel counts of the tcgen05 / TMEM / TMA SASS mnemonics in the shipped library.
usage: python tools/sass_evidence.py [path/to/libsegtran_b200.so] > profiles/r2_sass_tcgen05.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "segtran_b200/libsegtran_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
PAT = re.compile(r"\b(UTC[A-Z]*MMA(?:\.[\w.]+)?|LDTM(?:\.[\w.]+)?|STTM(?:\.[\w.]+)?|UTMALDG(?:\.[\w.]+)?|UTMASTG(?:\.[\w.]+)?|"
                 r"UTMAREDG(?:\.[\w.]+)?|UTCBAR(?:\.[\w.]+)?|UBLKCP(?:\.[\w.]+)?|HMMA(?:\.[\w.]+)?)\b")
per, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur:
        for tok in PAT.findall(line):
            per[cur][tok] += 1
print("# SASS evidence: tcgen05 / TMEM / TMA mnemonics in the shipped %s (sm_100a only)" % so)
print("# cuobjdump -sass | per-kernel count of UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG/UTMASTG/UTMAREDG (TMA load/store/reduce),")
print("# UTCBAR (tcgen05.commit), HMMA (legacy mma.sync: must be absent).  Regenerate with tools/sass_evidence.py")
print()
print("arch(s) in the fatbin:", ", ".join(archs))
tot = collections.Counter()
for k, c in per.items():
    if c:
        short = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+\d*", "", k)
        print("%-92s %s" % (short[-92:], "  ".join("%s x%d" % kv for kv in sorted(c.items()))))
        tot.update(c)
print()
print("totals:", "  ".join("%s x%d" % kv for kv in sorted(tot.items())))
print("kernels in the library: %d; kernels using the tensor core / TMA units: %d; HMMA (mma.sync) instructions: %d" % (
    len(per), sum(1 for c in per.values() if c), sum(v for k, v in tot.items() if k.startswith("HMMA"))))
