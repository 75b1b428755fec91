#!/usr/bin/env python
"""Scan every kernel of the built library for memory instructions that should not be there: generic LD / ST (a
shared- or global-memory pointer that lost its address space: slower, tracked on the long scoreboard) and LDL / STL
(local memory: spills, run-time indexed arrays).  Also counts the Tensor Memory and bulk-copy instructions.
usage: python scripts/sass_scan.py [library]   (needs cuobjdump; no GPU)"""
import collections, os, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rocoder_b200", "librocoder_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ("LD", "ST", "LDL", "STL", "LDS", "STS", "LDG", "STG", "LDTM", "STTM", "UBLKCP")
fn, counts = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and fn and m.group(1) in WATCH:
        counts[fn][m.group(1)] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("| kernel | LDS | STS | LDG | STG | LDTM | STTM | UBLKCP | generic LD | generic ST | LDL | STL |\n|---|---|---|---|---|---|---|---|---|---|---|---|")
bad = 0
for name, fn in sorted(zip(demangle, counts)):
    c = counts[fn]
    short = re.sub(r"\(.*", "", name).replace("rcg::", "").replace("(anonymous namespace)::", "")
    print("| %s | %s |" % (short, " | ".join(str(c[k]) for k in ("LDS", "STS", "LDG", "STG", "LDTM", "STTM", "UBLKCP", "LD", "ST", "LDL", "STL"))))
    bad += c["LD"] + c["ST"]
print("\ngeneric accesses in the library: %d" % bad)
