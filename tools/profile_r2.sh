#!/bin/bash
# round-2 ncu captures (run under gpurun, ONE GPU).  Each profiled command runs plainly first (&&), as the recipe
# requires; reports are exported to CSV on the box and the .ncu-rep files deleted (gpurun_out/ is capped at 64 MiB).
set -x
mkdir -p gpurun_out
OUT=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
case "$1" in
launches)
  $B > $OUT/r2p_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r2p_launches.csv $B > $OUT/r2p_ncu1.log 2>&1
  gzip -f $OUT/r2p_launches.csv ;;
tail)
  $B > $OUT/r2p_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"find_homography_kernel|pnp_ransac_kernel|hamming_top2_kernel" -s 9 -c 3 -o $OUT/r2p_tail $B > $OUT/r2p_ncu2.log 2>&1
  ncu -i $OUT/r2p_tail.ncu-rep --page raw --csv > $OUT/r2p_tail_raw.csv
  ncu -i $OUT/r2p_tail.ncu-rep --page source --csv > $OUT/r2p_tail_source.csv 2>/dev/null
  gzip -f $OUT/r2p_tail_raw.csv $OUT/r2p_tail_source.csv; rm -f $OUT/r2p_tail.ncu-rep ;;
tailsrc)
  # per-CUDA-line stall samples of the two latency-bound tail kernels (RANSAC homography, PnP-RANSAC)
  $B > $OUT/r2p_plain4.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"find_homography_kernel|pnp_ransac_kernel" -s 6 -c 2 -o $OUT/r2p_tailsrc $B > $OUT/r2p_ncu4.log 2>&1
  ncu -i $OUT/r2p_tailsrc.ncu-rep --page raw --csv > $OUT/r2p_tailsrc_raw.csv
  ncu -i $OUT/r2p_tailsrc.ncu-rep --page source --print-source cuda --csv > $OUT/r2p_tailsrc_cuda.csv 2>/dev/null
  gzip -f $OUT/r2p_tailsrc_raw.csv $OUT/r2p_tailsrc_cuda.csv; rm -f $OUT/r2p_tailsrc.ncu-rep ;;
extract)
  E="python bench.py --workload extract --steps 1 --warmup 3 --no-cpu-baseline"
  $E > $OUT/r2p_plain3.log 2>&1 &&
  ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats --section InstructionStats \
      --clock-control none -k regex:"k_fed|k_hessian|k_prep_level|k_extrema|k_mldb|k_orientation|k_contrast|k_gray|k_halfsample" -s 249 -c 83 -o $OUT/r2p_extract $E > $OUT/r2p_ncu3.log 2>&1
  ncu -i $OUT/r2p_extract.ncu-rep --page raw --csv > $OUT/r2p_extract_raw.csv
  gzip -f $OUT/r2p_extract_raw.csv; rm -f $OUT/r2p_extract.ncu-rep ;;
esac
du -sh $OUT; ls -la $OUT
