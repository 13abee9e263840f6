#!/usr/bin/env python
"""Summarise ncu reports into profiles/: key raw metrics per kernel + top stall lines of the source page.
usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [...] > profiles/rN_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_uniform.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_membar.ratio",
    "smsp__average_warp_latency_issue_stalled_sleeping.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
    "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
    "smsp__average_warp_latency_issue_stalled_selected.ratio", "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio",
]


def raw(rep):
    """One dict per kernel of the report."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [{h: (v, u) for h, u, v in zip(hdr, units, vals)} for vals in rows[2:] if vals]


STALLS = ["stall_barrier", "stall_branch_resolving", "stall_dispatch", "stall_drain", "stall_lg", "stall_long_sb",
          "stall_math", "stall_membar", "stall_mio", "stall_misc", "stall_no_inst", "stall_not_selected",
          "stall_selected", "stall_short_sb", "stall_sleep", "stall_tex", "stall_wait"]


def source(rep, top=16):
    """(stall totals over the kernel, top sampled SASS lines with their dominant stall reason)"""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next((i for i, r in enumerate(rows) if r and r[0] == "Address"), None)
    if hi is None:
        return {}, []
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    samp, src = ix.get("# Samples"), ix.get("Source")
    totals = {k: 0 for k in STALLS}
    items = []
    for r in rows[hi + 1:]:
        try:
            n = int(float(r[samp]))
        except (ValueError, IndexError, TypeError):
            continue
        st = {}
        for k in STALLS:
            try:
                st[k] = int(float(r[ix[k]]))
            except (ValueError, KeyError, IndexError):
                st[k] = 0
            totals[k] += st[k]
        items.append((n, r[src].strip(), max(st, key=st.get) if n else "-"))
    tot = sum(i[0] for i in items) or 1
    items.sort(key=lambda t: -t[0])
    return {k: v / tot for k, v in totals.items()}, [(n, n / tot, s, d) for n, s, d in items[:top]]


for rep in sys.argv[1:]:
    kernels = raw(rep)
    for ki, m in enumerate(kernels):
        name = m.get("Kernel Name", ("?", ""))[0]
        print("=" * 100)
        print(rep, "|", name[:90])
        for k in KEYS:
            if k in m:
                print(f"  {k:85s} {m[k][0]:>16s} {m[k][1]}")
        if ki > 0:
            continue                      # the source page of a multi-kernel report covers its first kernel only
        totals, src = source(rep)
        if totals:
            top = sorted(totals.items(), key=lambda kv: -kv[1])[:10]
            print("  -- warp-state samples by reason (share of all samples) --")
            print("  " + "  ".join(f"{k.replace('stall_', '')}={v * 100:.1f}%" for k, v in top))
            print("  -- top sampled SASS lines --")
            for n, share, line, dom in src:
                print(f"  {n:8d} {share * 100:5.1f}%  {dom.replace('stall_', ''):14s} {line[:90]}")
