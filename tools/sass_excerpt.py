#!/usr/bin/env python3
"""SASS evidence of the shipped conv kernel (runs on CPU: cuobjdump only reads the .so).

    python tools/sass_excerpt.py [mangled-name substring] > profiles/r2_sass_conv_tcp.txt

Default kernel: pe::conv_tcp_kernel<128, 2, 5, 1, 0> (the instantiation 312 of a step's 370 conv launches run).
Prints instruction counts of the Blackwell-native mnemonics and every tensor-core / TMA / TMEM / barrier instruction
in program order."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "caffe_rtpose_b200", "libposeengine.so")
want = sys.argv[1] if len(sys.argv) > 1 else "conv_tcp_kernelILi128ELi2ELi5ELi1ELi0EE"

out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
body, name, on = [], None, False
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        on = want in m.group(1)
        if on:
            name = m.group(1)
        continue
    if on and re.match(r"\s*/\*[0-9a-f]{4,6}\*/", line):
        body.append(line.split("/* 0x")[0].rstrip())
if not body:
    sys.exit("no function matching %r in %s" % (want, SO))
demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]

KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACMDFLUSH", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "ELECT", "HMMA", "STS", "LDS",
        "STG", "F2FP", "MEMBAR"]
cnt = collections.Counter()
for ln in body:
    op = re.sub(r"^@!?U?P[0-9T]\s+", "", ln.split("*/", 1)[1].strip()).split(" ")[0].split(".")[0]
    cnt[op] += 1
print("# SASS of the shipped conv kernel: %s (cuobjdump -sass libposeengine.so, sm_100a; tools/sass_excerpt.py)" % demangled)
print("# %d instructions.  Blackwell-native mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA load," % len(body))
print("# UTMASTG = TMA store (+ UTMACMDFLUSH = bulk commit_group), UTMAPF = TMA L2 prefetch, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit")
print("# (.MULTICAST = both CTAs of the pair), UTCATOMSWS = tcgen05.alloc/dealloc.  No HMMA (legacy mma.sync).  The STG instructions belong to")
print("# the planar fp32 store of the last stage and the direct-store fallback; the activation planes leave through STS.128 + UTMASTG.")
print("\n## instruction counts\n")
for k in KEYS:
    print("%-14s %d" % (k, cnt.get(k, 0)))
print("\n## every tensor-core / TMA / TMEM / barrier instruction, in program order\n")
pat = re.compile(r"\b(UTCHMMA|UTMALDG|UTMASTG|UTMAPF|UTMACMDFLUSH|LDTM|UTCBAR|UTCATOMSWS|SYNCS|ELECT|UCGABAR\w*|MEMBAR|FENCE)\b")
for ln in body:
    if pat.search(ln):
        print(ln)
