#!/bin/bash
# round-2 ncu captures (run under gpurun, ONE GPU): launch list of the default bench, --set full of the matcher, the
# RANSAC / PnP kernels and the extraction kernels.  Each command runs plainly first (&&), as the recipe requires.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --distinct 128"
$B > gpurun_out/r2p_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2p_launches.csv $B > gpurun_out/r2p_ncu1.log 2>&1
$B > gpurun_out/r2p_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"find_homography_kernel|pnp_ransac_kernel|hamming_top2_kernel" -s 9 -c 6 -o gpurun_out/r2p_tail $B > gpurun_out/r2p_ncu2.log 2>&1
E="python bench.py --workload extract --steps 1 --warmup 3 --no-cpu-baseline --extract-frames 64"
$E > gpurun_out/r2p_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_fed|k_hessian|k_prep_level|k_extrema|k_mldb|k_orientation|k_contrast|k_gray|k_halfsample" -s 300 -c 100 -o gpurun_out/r2p_extract $E > gpurun_out/r2p_ncu3.log 2>&1
ls -la gpurun_out/
