#!/bin/bash
# Round profile evidence (run under gpurun, one GPU): bench numbers first (never under ncu), then
# (a) launch list of a bench step, (b) --set full of the Hamming matcher, (c) of the tcgen05 L2 matcher,
# (d) of the extraction kernels.  Raw CSV pages are exported on the box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
set -x
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/final_bench_ref_n1.json 2> gpurun_out/final_bench_ref_n1.err
timeout 300 python bench.py --workload match --steps 5 --warmup 3 > gpurun_out/final_match_n1.json 2> gpurun_out/final_match_n1.err
timeout 200 python tools/bench_extract.py 256 64 3 > gpurun_out/final_extract.json 2> gpurun_out/final_extract.err
timeout 200 python tools/bench_match_l2.py 3163 2000000 64 > gpurun_out/final_l2_d64.json 2>&1
timeout 200 python tools/bench_match_l2.py 3163 2000000 128 > gpurun_out/final_l2_d128.json 2>&1
# (a) launch list
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 2 --warmup 3 --frames 16 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# (b) Hamming matcher
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hamming_top2 -s 3 -c 1 -o gpurun_out/prof_hamming_final -f \
  python tools/bench_match.py 3163 2000000 > gpurun_out/ncu_hamming.log 2>&1
ncu -i gpurun_out/prof_hamming_final.ncu-rep --page raw --csv > gpurun_out/prof_hamming_final_raw.csv 2>/dev/null
# (c) tcgen05 L2 matcher
timeout 400 ncu --set full --clock-control none --import-source on -k regex:l2_candidates -s 2 -c 2 -o gpurun_out/prof_l2_final -f \
  python tools/bench_match_l2.py 3163 2000000 64 > gpurun_out/ncu_l2.log 2>&1
ncu -i gpurun_out/prof_l2_final.ncu-rep --page raw --csv > gpurun_out/prof_l2_final_raw.csv 2>/dev/null
# (b2) DRAM traffic of the matcher launch at the bench's own shape (64 frames x 289 k rows): one cheap pass
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:hamming_top2 -s 4 -c 1 --csv \
  --log-file gpurun_out/traffic_hamming_pipeline.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
# (d) extraction kernels
bash tools/ncu_extract.sh > gpurun_out/ncu_extract_sh.log 2>&1
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -30
