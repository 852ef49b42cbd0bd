#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libssf.so (cuobjdump -sass), so that the Blackwell-native claims of DESIGN.md
(packed fp32x2 arithmetic, 256-bit gathers, bulk-copy TMA, 64-bit reductions, cluster barriers) can be checked
without rebuilding:  python tools/sass_counts.py > profiles/sass_counts_rN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "supersurfel_fusion_b200", "libssf.so")
WATCH = [("FFMA2", r"\bFFMA2\b"), ("FMUL2/FADD2", r"\b(FMUL2|FADD2)\b"), ("LDG.E.256", r"\bLDG\.E\.[A-Z0-9.]*?\b256\b"),
         ("LDG.E.128", r"\bLDG\.E\.[A-Z0-9.]*?\b128\b"), ("REDG.64", r"\bREDG?\.E\.[A-Z0-9.]*?\b64\b"), ("ATOMG", r"\bATOMG\b"),
         ("UBLKCP (TMA bulk)", r"\bUBLKCP\b"), ("UTMALDG (TMA tensor)", r"\bUTMALDG\b"), ("SYNCS (mbarrier)", r"\bSYNCS\b"),
         ("UCGABAR (cluster barrier)", r"\bUCGABAR"), ("SHFL", r"\bSHFL\b"), ("MATCH", r"\bMATCH\b"),
         ("DFMA", r"\bDFMA\b"), ("MUFU", r"\bMUFU\b"), ("UTCMMA/HMMA (tensor core)", r"\b(UTC[A-Z]*MMA|HMMA|IMMA)\b")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            cur = cur.replace("void ", "").replace("ssf::", "")
            kernels[cur] = collections.Counter()
            continue
        if cur and re.search(r"/\*[0-9a-f]{4,}\*/", line):
            kernels[cur]["instructions"] += 1
            for name, pat in WATCH:
                if re.search(pat, line):
                    kernels[cur][name] += 1
    print("# cuobjdump -sass %s   arch: %s   kernels: %d" % (os.path.relpath(LIB, ROOT), ",".join(arch), len(kernels)))
    total = collections.Counter()
    for k, c in kernels.items():
        total.update(c)
        tags = "  ".join("%s=%d" % (n, c[n]) for n, _ in WATCH if c[n])
        print("%-46s instr=%5d  %s" % (k[:46], c["instructions"], tags))
    print("# total: instr=%d  %s" % (total["instructions"], "  ".join("%s=%d" % (n, total[n]) for n, _ in WATCH)))


if __name__ == "__main__":
    main()
