# refreshes the round's bench / extraction numbers and the ncu --set full capture of the extraction kernels
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu 2>&1 | tail -2 > gpurun_out/final_pytest.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
timeout 300 python bench.py --workload extract > gpurun_out/bench_extract.json 2> gpurun_out/bench_extract.err
timeout 400 python bench.py --workload build > gpurun_out/bench_build.json 2> gpurun_out/bench_build.err
timeout 200 python tools/bench_extract.py 256 64 3 > gpurun_out/final_extract.json 2> gpurun_out/final_extract.err
bash tools/ncu_extract.sh > gpurun_out/ncu_extract_sh.log 2>&1
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/final_pytest.txt
tail -c 300 gpurun_out/final_bench_n1.json
