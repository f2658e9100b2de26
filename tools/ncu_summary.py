#!/usr/bin/env python3
"""Text summary of one kernel of an .ncu-rep (raw page + per-opcode stall samples of the source page):
    python tools/ncu_summary.py gpurun_out/x/prof.ncu-rep "header line" > profiles/rNN_ncu_<kernel>.txt
Runs here (no GPU needed): ncu -i ... --page raw/source --csv."""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    for line in sys.argv[2:]:
        print(line)
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"{'Kernel Name':<112} {vals[col['Kernel Name']]}")
    for k in KEYS:
        if k in col:
            print(f"{k:<96} {units[col[k]]:<16} {vals[col[k]]}")
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            print(f"{h:<96} {units[col[h]]:<16} {vals[col[h]]}")
    src = page(rep, "source")
    if len(src) > 2:
        shdr, data = src[1], src[2:]
        ix = {h: i for i, h in enumerate(shdr)}
        stalls = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
        ex, samp, per = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
        for r in data:
            toks = r[ix["Source"]].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            ex[op] += int(r[ix["Instructions Executed"]])
            samp[op] += int(r[ix["# Samples"]])
            for c in stalls:
                per[op][c[6:]] += int(r[ix[c]])
        total = sum(samp.values()) or 1
        print("\nwarp-stall samples by opcode (source page; top 14), with the opcode's executed warp instructions:")
        for op, s in samp.most_common(14):
            top = ", ".join(f"{k}={v}" for k, v in per[op].most_common(3))
            print(f"  {op:<30} executed {ex[op]:>12}  samples {s:>7} ({100 * s / total:4.1f} %)  {top}")
        alls = collections.Counter()
        for op in per:
            alls.update(per[op])
        print("  all samples by reason:", dict(alls.most_common(9)))


if __name__ == "__main__":
    main()
