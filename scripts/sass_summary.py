#!/usr/bin/env python
"""profiles/<tag>_sass.md: per kernel of libmpgnn_b200.so, the SASS mnemonics that prove the tensor-core / TMA path
(UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA load/store, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit)."""
import collections
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = "mpgnn-metapath-graph-neural-network_b200/libmpgnn_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTMALDG(?:\.\dD)?|UTMASTG(?:\.\dD)?|LDTM|STTM|UTCBAR|UTMAPF|SYNCS)\b")
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur:
        for hit in pat.findall(line):
            counts[cur][hit.split(".")[0] + (".2CTA" if hit.endswith(".2CTA") else "")] += 1
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR"]
out = ["# SASS evidence `%s`" % tag, "",
       "`cuobjdump -sass %s` (sm_100a), instruction counts per kernel; kernels without any of these are plain SIMT." % lib, "",
       "| kernel | " + " | ".join(cols) + " |", "|---|" + "---:|" * len(cols)]
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    if any(c[x] for x in cols):
        out.append("| `%s` | " % k[:90] + " | ".join(str(c[x]) for x in cols) + " |")
out.append("| **all %d kernels** | " % len(counts) + " | ".join(str(tot[x]) for x in cols) + " |")
open("profiles/%s_sass.md" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out))
