"""Writes profiles/r2_summary.md from the committed ncu exports (profiles/r2_*.csv.gz): launch-list shares of the bench
command, the per-kernel tables of tools/summarize_ncu.py and the matcher's pipe metrics.
    python tools/make_r2_summary.py > profiles/r2_summary.md"""
import collections
import csv
import gzip
import io
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
P = lambda f: os.path.join(ROOT, "profiles", f)

rows = [r for r in csv.reader(io.TextIOWrapper(gzip.open(P("r2_launches.csv.gz")))) if len(r) > 14 and r[0].isdigit()]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("dunk::<unnamed>::", "").replace("void ", "")
    tot[name] += float(r[14])
    cnt[name] += 1
T = sum(tot.values())
print("# Round 2 — ncu evidence for the final kernels\n")
print("## Launch list of the bench command\n")
print("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`")
print(f"(`tools/profile_r2.sh launches`; {len(rows)} launches: DB build of 85 tiles, warp of the 512 query frames, warm-up + timed + e2e + quality +")
print("profiled steps of 256 frames; raw list: `profiles/r2_launches.csv.gz`).  Per-launch times under ncu are cold-cache and serialised: compare SHARES,")
print("not absolutes.  The library's own CUDA-event profile of a bench step gives `match.hamming_top2` 68 % (`profiles/r2_bench_n1.json`:")
print("`roofline.share_of_step`); the list below also holds the set-up kernels.\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for name, v in tot.most_common(24):
    print(f"| `{name}` | {cnt[name]} | {v / 1e6:.3f} | {100 * v / T:.1f} % |")
print()
for f, title in (("r2_tail_raw.csv.gz", "`--set full` of the matcher and the two tail kernels inside the bench (`tools/profile_r2.sh tail`)"),
                 ("r2_extract_raw.csv.gz", "The 83 scale-space / detector / descriptor launches of ONE 256-frame step of `bench.py --workload extract` "
                                           "(`tools/profile_r2.sh extract`: `-s 249 -c 83`, sum 23.1 ms = the step)")):
    print("## " + title + "\n")
    print(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_ncu.py"), P(f)], capture_output=True, text=True).stdout)
print("""Every scale-space kernel is instruction-issue bound (issue-active 68-80 %: `k_gray_gauss9_reg` 80, `k_contrast_modg_reg` 77, `k_prep_level_reg` 76,
`k_fed_reg` 71-72, `k_hessian_reg` 68-72), which is why their HBM fractions in `stages_ms_per_step` stop at 0.28-0.63; `k_contrast_hist` (0.85) and
`k_halfsample_x2` (0.98) are the two that are bandwidth-bound.  `k_mldb` / `k_orientation` are L1TEX / L2 gather bound (long- and short-scoreboard, MIO throttle).

## Matcher pipes (`profiles/r2_tail_raw.csv.gz`, `hamming_top2_kernel<4>`, 133.6 k queries x 152 750 rows, 44.2 ms under ncu)

| metric | value |
|---|---|
| `sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active` | 82.8 % |
| `sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active` (POPC) | 81.1 % |
| `sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active` | 6.1 % |
| `smsp__issue_active.avg.pct_of_peak_sustained_active` | 60.2 % |
| `dram__bytes_read.sum` / `dram__bytes_write.sum` | 18.35 MB / 0.001 MB (algorithmic: 9.8 MB of DB rows + 8.6 MB of query rows) |
| `smsp__inst_executed.sum` | 30.6 G warp instructions = 48.0 per (query, row) pair and lane (SASS: 29 LOP3 + 8 POPC + 3 IMAD + 1.3 IADD3 + 1.1 ISETP + 0.25 LDS.128 in the hot loop) |

## Tail kernels

`profiles/r2_tail_phase_times_before.txt` / `_after.txt`: `clock64` stamps of block 0 (build `make -C cubesat-apds_b200/csrc timing`, `tools/phase_times.py`) before and after
the 128-thread CTAs; `profiles/r2_tail_microbench.jsonl`: `tools/bench_tail.py` at cv2's probe shapes (N = 1000, 40 % outliers).

## tcgen05 float matcher

`profiles/r2_match_l2_tcgen05.jsonl` (`tools/bench_match_l2.py`): 3 163 x 2 M rows D = 64 2.10 ms (main stage 499 TFLOP/s tf32), D = 128 3.20 ms (635 TFLOP/s),
16 384 x 2 M rows D = 64 9.7 ms, 256 x 2 M rows 0.66 ms.
""")
