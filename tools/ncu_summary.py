#!/usr/bin/env python
"""Text summary of one kernel of an .ncu-rep (read here, without a GPU): python tools/ncu_summary.py rep.ncu-rep "note" > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("kernel:", d.get("Kernel Name", "?"))
        if note:
            print("note:", note)
        for k in KEEP:
            if k in d:
                print(f"  {k} {d[k]} {units[hdr.index(k)]}")
        print("  stall cycles per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
        for h, v in zip(hdr, vals):
            if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
                try:
                    if float(v) > 0.02:
                        print("    ", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
                except ValueError:
                    pass
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 3:
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        data = [r for r in rows[2:] if len(r) == len(hdr)]
        tot = sum(int(r[ix["# Samples"]]) for r in data)
        mx = max(int(r[ix["Instructions Executed"]]) for r in data)
        hot = [r for r in data if int(r[ix["Instructions Executed"]]) > 0.2 * mx]
        print(f"  hot loop: {len(hot)} SASS instructions executed > 20 % of the maximum count, "
              f"{sum(int(r[ix['# Samples']]) for r in hot) / max(tot, 1):.1%} of all stall samples")
        ops = {}
        for r in hot:
            op = r[ix["Source"]].split()[0]
            if op.startswith("@"):
                op = r[ix["Source"]].split()[1]
            ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
        print("  hot loop opcode mix:", ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
        print("  top stall sites (share of all samples):")
        for r in sorted(hot, key=lambda r: -int(r[ix["# Samples"]]))[:12]:
            print(f"    {int(r[ix['# Samples']]) / max(tot, 1):6.2%}  {r[ix['Source']].strip()[:70]}")


if __name__ == "__main__":
    main()
