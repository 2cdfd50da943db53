"""Summarise an ncu report of the refine kernel (run here, no GPU needed):

    python profiles/summarize_ncu.py gpurun_out/prof_r03.ncu-rep profiles/r01_refine_summary.md

Extracts the launch metrics the roofline discussion uses (duration, registers, occupancy, issue
utilisation, stall reasons, DRAM bytes) and, from the source page, the share of executed
instructions per phase of csrc/ctk_solver.cuh (needs -lineinfo, which the build always passes).
"""
import bisect
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]
PHASES = [
    ("setup_variables", r"CTK_DEV_BIG int setup_variables"),
    ("build: box / tables", r"CTK_DEV_BIG int build_pixels"),
    ("build: walk box, compact union", r"// walk the box in C order"),
    ("build: per-feature and shared-pixel lists", r"// Per-feature pixel lists"),
    ("load_features / feat", r"CTK_DEV void load_features"),
    ("geometry / model value / derivatives", r"CTK_DEV Geo geometry"),
    ("evaluate", r"CTK_DEV_BIG double evaluate"),
    ("accumulate: features", r"CTK_DEV_BIG void accumulate"),
    ("accumulate: shared pixels (cross blocks)", r"// cross blocks over the pixels"),
    ("constraint views", r"// ---- distance constraints as augmented"),
    ("solve: assemble / scale / freeze", r"CTK_DEV_BIG bool solve"),
    ("predicted", r"CTK_DEV_BIG double predicted"),
    ("minimise (LM loop)", r"CTK_DEV_BIG int minimise"),
    ("run (outer loop, write-back)", r"CTK_DEV void run"),
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep, out, solver_path="clustertracking_b200/csrc/ctk_solver.cuh"):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu summary of `%s`" % rep, "",
             "Captured with `ncu --set full --clock-control none --import-source on` under gpurun "
             "(profiled launches are cold-cache and serialised: compare shares, not absolutes).", ""]
    for k, row in enumerate(data):
        lines += ["## launch %d: `%s`" % (k, row[idx["Kernel Name"]][:110]), "",
                  "| metric | value | unit |", "|---|---|---|"]
        for m in METRICS:
            if m in idx:
                lines.append("| %s | %s | %s |" % (m, row[idx[m]], units[idx[m]]))
        lines.append("")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv",
                                           "--print-source", "cuda,sass"]))))
    start = None
    for i, r in enumerate(src):
        if len(r) >= 2 and r[0] == "File Path" and "ctk_solver.cuh" in r[1]:
            start = i
    if start is not None:
        text = open(solver_path).read().splitlines()
        marks = [(0, "inlined helpers (reductions, unpacking, exp, sync)")]
        for name, pat in PHASES:
            for n, line in enumerate(text, 1):
                if re.search(pat, line):
                    marks.append((n, name))
                    break
        marks.sort()
        keys = [m[0] for m in marks]
        h = src[start + 2]
        ci = {c: k for k, c in enumerate(h)}
        agg = {}
        for r in src[start + 3:]:
            if len(r) < 10:
                break
            try:
                ln = int(r[0])
                inst = int(r[ci["Instructions Executed"]])
                smp = int(r[ci["# Samples"]])
                tin = int(r[ci["Thread Instructions Executed"]])
            except ValueError:
                continue
            name = marks[bisect.bisect_right(keys, ln) - 1][1]
            a = agg.setdefault(name, [0, 0, 0])
            a[0] += inst
            a[1] += smp
            a[2] += tin
        tot = sum(a[0] for a in agg.values()) or 1
        tots = sum(a[1] for a in agg.values()) or 1
        lines += ["## executed warp instructions per phase (first profiled launch)", "",
                  "Line numbers resolved against the CURRENT `%s`; regenerate right after capturing."
                  % solver_path, "",
                  "| phase | instructions % | samples % | active lanes / instruction |", "|---|---|---|---|"]
        for _, name in marks:
            if name in agg:
                a = agg[name]
                lines.append("| %s | %.1f | %.1f | %.1f |" % (name, 100. * a[0] / tot, 100. * a[1] / tots,
                                                              a[2] / max(a[0], 1)))
        lines.append("")
    with open(out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main(*sys.argv[1:])
