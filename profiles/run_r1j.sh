set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_j.log 2>&1; tail -5 gpurun_out/pytest_j.log
python bench.py --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; tail -c 600 gpurun_out/bench_j.err
