#!/usr/bin/env python3
"""Per-kernel counts of the SASS mnemonics that prove the TMA / mbarrier / warp-collective path (UBLKCP = cp.async.bulk,
SYNCS = mbarrier, REDUX, SHFL, LDS) in the built library -> profiles/r2_sass_tma_evidence.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ds2i_b200", "lib", "libds2i_gpu.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.splitlines()
keys = ("UBLKCP", "SYNCS", "REDUX", "SHFL", "LDS")
cur, per = None, collections.OrderedDict()
for line in sass:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]+\*/", line):
        per[cur]["insts"] += 1
        for k in keys:
            if k in line:
                per[cur][k] += 1
out = ["# SASS evidence of the TMA path (round 2): `cuobjdump -sass ds2i_b200/lib/libds2i_gpu.so` of the committed sources, built by ds2i_b200/build.py",
       "# (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a).  UBLKCP = cp.async.bulk (1-D TMA bulk copy global -> shared), SYNCS = mbarrier ops",
       "# (SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK.TRANS64.TRYWAIT), REDUX = warp reductions, SHFL = warp shuffles.  Regenerate: python tools/sass_evidence.py",
       "", "%-110s %7s %7s %6s %6s %6s %6s" % ("kernel", "insts", "UBLKCP", "SYNCS", "REDUX", "SHFL", "LDS")]
tot = collections.Counter()
for k, d in per.items():
    out.append("%-110s %7d %7d %6d %6d %6d %6d" % (k[:110], d["insts"], d["UBLKCP"], d["SYNCS"], d["REDUX"], d["SHFL"], d["LDS"]))
    tot.update(d)
out.append("%-110s %7d %7d %6d %6d %6d %6d" % ("TOTAL", tot["insts"], tot["UBLKCP"], tot["SYNCS"], tot["REDUX"], tot["SHFL"], tot["LDS"]))
out += ["", "# first occurrences:"] + [l.rstrip() for l in sass if "UBLKCP" in l or "SYNCS" in l][:12]
open(os.path.join(ROOT, "profiles", "r2_sass_tma_evidence.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-16:]))
