"""Summarise an `ncu --page raw --csv` export: one row per captured launch with the metrics the
roofline discussion needs.  Usage: python tools/ncu_summary.py file.csv [more.csv ...]"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "us"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_reg"),
    ("launch__occupancy_limit_shared_mem", "occ_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("dram__bytes_read.sum", "dramR_MB"),
    ("dram__bytes_write.sum", "dramW_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_MB"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
    ("smsp__inst_executed.sum", "inst"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
]


def to_float(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}")
    print("| kernel | " + " | ".join(k for _, k in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in rows[hdr_i + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]].split("(")[0].split("::")[-1]
        out = []
        for m, short in KEYS:
            if m not in col:
                out.append("-")
                continue
            v = to_float(r[col[m]])
            u = units[col[m]]
            if v is None:
                out.append(r[col[m]])
                continue
            if short == "us":
                v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
            if short.endswith("_MB"):
                v = {"byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(u, v)
            out.append(f"{v:.3g}" if abs(v) < 1e6 else f"{v:.4g}")
        print(f"| {name} | " + " | ".join(out) + " |")
