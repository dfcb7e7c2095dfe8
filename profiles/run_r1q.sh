set -x
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r.log 2>&1; tail -3 gpurun_out/pytest_q.log
timeout 120 python bench.py --no-cpu-baseline --no-torch-gpu-baseline --no-other-configs > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo rc=$?; tail -c 300 gpurun_out/bench_r.err
