#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into the text files committed under profiles/.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep UNITS_PER_LAUNCH [out.txt]

UNITS_PER_LAUNCH = (site, chain) pairs per launch / 32, i.e. the divisor that turns warp-level
executed-instruction counts into "instructions per site-chain".
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "launch__block_size", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def ncu(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    out = open(sys.argv[3], "w") if len(sys.argv) > 3 else sys.stdout
    raw = ncu(rep, "raw")
    hdr, unit, val = raw[0], raw[1], raw[2]
    d = {h: (v, u) for h, u, v in zip(hdr, unit, val)}
    print(f"# {rep}\n# kernel: {d.get('Kernel Name', ('?',))[0]}", file=out)
    for k in KEYS:
        if k in d:
            print(f"{k:85s} {d[k][0]:>16s} {d[k][1]}", file=out)
    src = ncu(rep, "source", ["--print-source", "sass"])
    h = src[1]
    ia, isrc = h.index("Instructions Executed"), h.index("Source")
    byop, tot = collections.Counter(), 0
    for r in src[2:]:
        if len(r) <= ia:
            continue
        n = int(r[ia])
        parts = r[isrc].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        byop[op.split(".")[0].rstrip(";")] += n
        tot += n
    print(f"\n# executed warp-instructions: {tot}  = {tot / units:.1f} per (site, chain)", file=out)
    print("# opcode         per (site,chain)    share", file=out)
    for op, n in byop.most_common(24):
        print(f"{op:12s} {n / units:12.2f} {100 * n / tot:10.1f}%", file=out)


if __name__ == "__main__":
    main()
