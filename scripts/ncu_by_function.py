#!/usr/bin/env python
"""Per-function (of rcg_frame.h) / per-file breakdown of an ncu capture: instruction share, stall
samples, long-scoreboard and barrier samples, excess shared-memory wavefronts."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout)))
lines = open('rocoder_b200/csrc/rcg_frame.h').read().split('\n')
func_of, cur = {}, '?'
for i, l in enumerate(lines, start=1):
    m = re.match(r'RCG_HD [\w<>\s\*&:]*?(\w+)\(', l)
    if m: cur = m.group(1)
    func_of[i] = cur
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0]); curf = None; idx = None
for r in src:
    if not r: continue
    if r[0] == "File Path": curf = r[1].split("/")[-1]
    elif r[0] == "Line No": idx = {h: i for i, h in enumerate(r)}
    elif r[0].isdigit() and idx:
        def g(n):
            try: return float(r[idx[n]])
            except Exception: return 0.0
        key = func_of.get(int(r[0]), '?') if curf == 'rcg_frame.h' else curf
        a = agg[key]
        a[0] += g("Instructions Executed"); a[1] += g("Warp Stall Sampling (All Samples)"); a[2] += g("stall_long_sb"); a[3] += g("stall_barrier"); a[4] += g("L1 Wavefronts Shared Excessive")
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("| function / file | instr % | stall samples % | long_sb % | barrier % | excess smem wavefronts |\n|---|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[0] / ti < 0.002 and v[1] / ts < 0.002: continue
    print("| %s | %.1f | %.1f | %.1f | %.1f | %.3g |" % (k, 100 * v[0] / ti, 100 * v[1] / ts, 100 * v[2] / ts, 100 * v[3] / ts, v[4]))
