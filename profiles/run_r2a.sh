#!/bin/bash
# round 2, first GPU pass: full -m gpu suite (incl. the reference's own suite + CLI through the swap), bench N=1 as the driver runs it
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt
( time python -m pytest tests -m gpu -q -x --timeout 3000 ) > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; wc -c gpurun_out/r2a_bench.json; tail -c 1500 gpurun_out/r2a_bench.err
cp profiles/bench_last_full.json gpurun_out/r2a_bench_full.json 2>/dev/null
( time python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 ) > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
echo "ref rc=$?"; cat gpurun_out/r2a_bench_ref.json | head -c 1500
