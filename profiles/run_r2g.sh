#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"; wc -c gpurun_out/r2g_bench.json; tail -c 600 gpurun_out/r2g_bench.err
cp profiles/bench_last_full.json gpurun_out/r2g_bench_full.json 2>/dev/null
cat gpurun_out/r2g_bench.json
