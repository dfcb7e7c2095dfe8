set -x
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "graph" > gpurun_out/pytest_z.log 2>&1; tail -2 gpurun_out/pytest_z.log
timeout 150 python bench.py --no-cpu-baseline --no-torch-gpu-baseline --no-other-configs --no-variants > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err; echo rc=$?; tail -c 300 gpurun_out/bench_z.err
