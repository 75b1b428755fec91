#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel capture made with --set full --import-source on) into a small
markdown file under profiles/.  usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.md [note]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = [f"# ncu summary: {rep.split('/')[-1]}", "", note, "",
             f"kernel: `{m.get('Kernel Name', ('?', ''))[0]}`", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in m:
            lines.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
    stalls = sorted(((float(v[0]), h) for h, v in m.items()
                     if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v[0]),
                    reverse=True)
    lines += ["", "## warp stall reasons (warps per issue-active cycle)", ""]
    for v, h in stalls[:8]:
        lines.append(f"- {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
    # per source line
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    agg, cur, idx = {}, None, None
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            idx = {h: i for i, h in enumerate(r)}
        elif r[0].isdigit() and idx:
            def g(name):
                try:
                    return float(r[idx[name]])
                except Exception:
                    return 0.0
            agg[(cur, int(r[0]))] = (g("Instructions Executed"), g("Warp Stall Sampling (All Samples)"), r[1][:100],
                                     g("L1 Wavefronts Shared Excessive"))
    ti = sum(v[0] for v in agg.values()) or 1.0
    ts = sum(v[1] for v in agg.values()) or 1.0
    lines += ["", "## hottest source lines (share of executed warp instructions / of stall samples)", "",
              "| file:line | instr % | samples % | excess smem wavefronts | source |", "|---|---|---|---|---|"]
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
        lines.append(f"| {f}:{l} | {100 * v[0] / ti:.2f} | {100 * v[1] / ts:.2f} | {v[3]:.3g} | `{v[2].strip()}` |")
    pf, ps = collections.Counter(), collections.Counter()
    for (f, l), v in agg.items():
        pf[f] += v[0]
        ps[f] += v[1]
    lines += ["", "## by file", ""]
    for f in pf:
        lines.append(f"- {f}: {100 * pf[f] / ti:.1f}% of instructions, {100 * ps[f] / ts:.1f}% of stall samples")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
