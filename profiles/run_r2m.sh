#!/bin/bash
# multi-GPU bench as the driver launches it
N=$1
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc=$?"; wc -c gpurun_out/r2_bench_n$N.json; tail -c 500 gpurun_out/r2_bench_n$N.err; cat gpurun_out/r2_bench_n$N.json
cp profiles/bench_last_full.json gpurun_out/r2_bench_n${N}_full.json 2>/dev/null
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
