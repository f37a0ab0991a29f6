#!/bin/bash
# multi-GPU runs on one box: sharded matcher (config 3) at N = 8, 4, 2 and the pipeline (config 5) at N = 4, 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # n workload tag
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus $1 --steps 5 --warmup 3 --workload $2 > gpurun_out/scale_$3_n$1.json 2> gpurun_out/scale_$3_n$1.err
  echo "rc=$? $3 n=$1: $(tail -c 400 gpurun_out/scale_$3_n$1.json | head -c 400)"
}
for n in 8 4 2; do run $n match match; done
for n in 4 8; do run $n pipeline pipe; done
