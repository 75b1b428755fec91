#!/usr/bin/env python
"""Per-function breakdown of an ncu capture over every header of rocoder_b200/csrc (ncu_by_function.py knows
rcg_frame.h only): share of executed instructions, of stall samples, and of the long-scoreboard / barrier / lg samples.
usage: python scripts/ncu_by_function_any.py report.ncu-rep"""
import collections, csv, glob, io, os, re, subprocess, sys
rep = sys.argv[1]
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout)))
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rocoder_b200", "csrc")
def funcs(path):
    m, cur = {}, "?"
    for i, l in enumerate(open(path).read().split("\n"), start=1):
        g = re.match(r"(?:template.*>\s*)?RCG_HD [\w<>\s\*&:]*?(\w+)\(", l) or re.match(r"\s*__device__ .*?(\w+)\(", l)
        if g: cur = g.group(1)
        m[i] = cur
    return m
fm = {os.path.basename(f): funcs(f) for f in glob.glob(os.path.join(root, "rcg_frame*.h")) + glob.glob(os.path.join(root, "rcg_*body.h"))}
cur, idx, agg = None, None, collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for r in rows:
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]
    elif r[0] == "Line No": idx = {h: i for i, h in enumerate(r)}
    elif r[0].isdigit() and idx and len(r) > 5:
        def g(n):
            try: return float(r[idx[n]])
            except Exception: return 0.0
        key = fm[cur].get(int(r[0]), "?") if cur in fm else cur
        a = agg[key]
        a[0] += g("Instructions Executed"); a[1] += g("Warp Stall Sampling (All Samples)"); a[2] += g("stall_long_sb"); a[3] += g("stall_barrier"); a[4] += g("stall_lg")
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("| function / file | instr % | stall samples % | long_sb % | barrier % | lg % |\n|---|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[0] / ti < 0.003 and v[1] / ts < 0.003: continue
    print("| %s | %.1f | %.1f | %.1f | %.1f | %.1f |" % (k, 100 * v[0] / ti, 100 * v[1] / ts, 100 * v[2] / ts, 100 * v[3] / ts, 100 * v[4] / ts))
