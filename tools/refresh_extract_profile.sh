cd /root/repo
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
timeout 200 python tools/bench_extract.py 256 64 3 > gpurun_out/final_extract.json 2> gpurun_out/final_extract.err
bash tools/ncu_extract.sh > gpurun_out/ncu_extract_sh.log 2>&1
rm -f gpurun_out/*.ncu-rep
tail -c 300 gpurun_out/final_bench_n1.json
