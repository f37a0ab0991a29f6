#!/bin/bash
# ncu --set full captures of the extraction kernels (16-frame batch keeps ncu's save/restore cheap);
# raw CSV pages are exported on the box, the .ncu-rep files are kept only if small
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_fed|k_hessian|k_prep_level|k_gray_gauss9|k_contrast_modg" --launch-skip 120 -c 17 \
  -o gpurun_out/prof_scale -f python tools/bench_extract.py 16 16 1 > gpurun_out/ncu_scale.log 2>&1
ncu -i gpurun_out/prof_scale.ncu-rep --page raw --csv > gpurun_out/prof_scale_raw.csv 2>/dev/null
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_mldb|k_orientation|k_extrema" --launch-skip 36 -c 18 \
  -o gpurun_out/prof_desc -f python tools/bench_extract.py 16 16 1 > gpurun_out/ncu_desc.log 2>&1
ncu -i gpurun_out/prof_desc.ncu-rep --page raw --csv > gpurun_out/prof_desc_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/prof_scale.ncu-rep gpurun_out/prof_desc.ncu-rep; do
  if [ $(stat -c %s $f) -gt 25000000 ]; then rm -f $f; fi
done
