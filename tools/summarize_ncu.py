#!/usr/bin/env python
"""Turns `ncu --page raw --csv` exports (gzip or plain) into the markdown tables kept under profiles/:
one row per kernel class (launches aggregated by name): launches, total / mean duration, share of the captured time,
registers, achieved occupancy, issue-active %, DRAM bytes per launch and throughput, top stall reasons.
    python tools/summarize_ncu.py gpurun_out/r2p_extract_raw.csv.gz [more.csv.gz ...] > profiles/r2_extract_ncu.md"""
import csv
import gzip
import io
import re
import sys
from collections import OrderedDict


def load(path):
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = [r for r in csv.reader(f) if r]
    # the log may start with ncu's ==PROF== lines
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[start], rows[start + 1], rows[start + 2:]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def short(name):
    name = name.replace("void ", "")
    name = re.sub(r"\(.*", "", name)                      # drop the argument list
    return name.split("::")[-1].strip()


def main(paths):
    for path in paths:
        hdr, units, data = load(path)
        ix = {h: i for i, h in enumerate(hdr)}

        def col(r, key):
            return num(r[ix[key]]) if key in ix else None
        groups = OrderedDict()
        for r in data:
            groups.setdefault(short(r[ix["Kernel Name"]]), []).append(r)
        dur_unit = units[ix["gpu__time_duration.sum"]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(dur_unit, 1.0)       # -> microseconds
        total = sum(col(r, "gpu__time_duration.sum") * scale for r in data)
        stall_cols = [h for h in hdr if "average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
        print(f"### {path.split('/')[-1]} — {len(data)} launches, {total / 1e3:.3f} ms captured\n")
        print("| kernel | launches | total µs | share | mean µs | regs | warps active % | issue active % | DRAM MB / launch | DRAM GB/s | top stalls (warps per issue) |")
        print("|---|---|---|---|---|---|---|---|---|---|---|")
        for name, rs in sorted(groups.items(), key=lambda kv: -sum(col(r, "gpu__time_duration.sum") for r in kv[1])):
            t = sum(col(r, "gpu__time_duration.sum") * scale for r in rs)

            def mean(key):
                v = [col(r, key) for r in rs if col(r, key) is not None]
                return sum(v) / len(v) if v else None

            def fmt(v, f="%.1f"):
                return "—" if v is None else f % v
            rd, wr = mean("dram__bytes_read.sum"), mean("dram__bytes_write.sum")
            bu = units[ix["dram__bytes_read.sum"]] if "dram__bytes_read.sum" in ix else "byte"
            bscale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(bu, 1e-6)
            mb = None if rd is None else (rd + (wr or 0)) * bscale
            gbs = None if mb is None else mb * 1e-3 / (t / len(rs) * 1e-6)
            stalls = []
            for h in stall_cols:
                v = mean(h)
                if v:
                    stalls.append((h.split("stalled_")[1].split("_per")[0], v))
            top = ", ".join(f"{k} {v:.2f}" for k, v in sorted(stalls, key=lambda x: -x[1])[:4] if k != "selected")
            print(f"| `{name}` | {len(rs)} | {t:.1f} | {100 * t / total:.1f} % | {t / len(rs):.1f} | {fmt(mean('launch__registers_per_thread'), '%.0f')} | "
                  f"{fmt(mean('sm__warps_active.avg.pct_of_peak_sustained_active'))} | {fmt(mean('smsp__issue_active.avg.pct_of_peak_sustained_active'))} | "
                  f"{fmt(mb, '%.2f')} | {fmt(gbs, '%.0f')} | {top} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1:])
