#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM (tcgen05.ld), UTMALDG (TMA tensor loads), UBLKCP (cp.async.bulk), SYNCS (mbarrier).   python tools/sass_summary.py [lib.so]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pnode_b200", "csrc", "libpnode_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|UTCATOMSWS|SYNCS|HMMA|DMMA|ATOMG|RED)\b")
counts, fn = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn:
        m = pat.search(line)
        if m:
            counts[fn][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# %s" % os.path.relpath(lib, ROOT))
print("# kernel | " + "mnemonic:count ...")
for mangled, name in zip(counts, names):
    c = counts[mangled]
    if not c:
        continue
    short = re.sub(r"\(.*", "", name).replace("pnode::", "").replace("void ", "")
    print("%-58s %s" % (short[:58], "  ".join("%s:%d" % kv for kv in sorted(c.items()))))
