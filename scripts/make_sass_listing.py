"""profiles/rNN_sass_listing.txt: per-kernel counts of the SASS instructions that prove the Blackwell-native paths
(tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UTMASTG/UBLKCP, packed fp32x2, FMNMX3, 256-bit stores).

    python scripts/make_sass_listing.py profiles/r02_sass_listing.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "dense2sparse-vit_b200", "libd2s_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA[A-Z0-9.]*|LDTM[A-Z0-9.]*|STTM[A-Z0-9.]*|UTMALDG[A-Z0-9.]*|UTMASTG[A-Z0-9.]*|UBLKCP[A-Z0-9.]*|HMMA[A-Z0-9.]*|"
                 r"FFMA2|FMUL2|FADD2|FMNMX3|STG\.E\.ENL2\.256|MUFU\.EX2|UTCBAR[A-Z0-9.]*|SYNCS[A-Z0-9.]*)\b")
fn, counts = None, collections.defaultdict(collections.Counter)
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        continue
    if fn is None:
        continue
    for t in pat.findall(ln):
        counts[fn][t if t.startswith(("STG", "MUFU")) else t.split(".")[0]] += 1
out = ["# SASS evidence (cuobjdump -sass dense2sparse-vit_b200/libd2s_b200.so, sm_100a only), instruction counts per kernel",
       "# UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = 1-D bulk copy,",
       "# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, FFMA2 / FMUL2 / FADD2 = packed fp32x2, HMMA = legacy mma.sync (none expected)", ""]
tot, allc = collections.Counter(), collections.Counter()
for fn in sorted(counts):
    c = counts[fn]
    allc.update(c)
    if not any(k.startswith(("UTC", "LDTM", "STTM", "UTMA", "UBLKCP", "HMMA")) for k in c):
        continue
    dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip()[:140]
    out += [dem, "    " + "  ".join(f"{k}={v}" for k, v in sorted(c.items()))]
    tot.update(c)
out += ["", "TOTAL (kernels above): " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items())),
        "TOTAL (whole library): " + "  ".join(f"{k}={v}" for k, v in sorted(allc.items())),
        "arch list of the cubins: " + " ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass))))]
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_listing.txt")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[-3:]))
