"""Turns an ncu report into the markdown summary kept under profiles/.
Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep "title / command" > profiles/x.md"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "sm__cycles_elapsed.avg", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none; per-launch, cold cache, serialised)\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## {d['Kernel Name']}\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEEP:
            if k in d:
                print(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
        stalls = []
        for k in hdr:
            if "pcsamp_warps_issue_stalled_" in k and not k.endswith("_not_issued"):
                try:
                    stalls.append((float(d[k]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        tot = sum(v for v, _ in stalls) or 1.0
        print("| warp stall samples (top) | " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in stalls[:7]) + " | |")
        print()


if __name__ == "__main__":
    main()
