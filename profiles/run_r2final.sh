#!/bin/bash
# final pass of the round: the whole parity file (incl. full-size and fuzz cases), then bench.py exactly as the driver runs it
mkdir -p gpurun_out
( time timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2final_pytest.log
( time timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --full-out gpurun_out/r2final_bench_full.json ) > gpurun_out/r2final_bench_line.json 2> gpurun_out/r2final_bench.err
echo "bench rc=$?"; wc -c gpurun_out/r2final_bench_line.json; tail -c 300 gpurun_out/r2final_bench.err
