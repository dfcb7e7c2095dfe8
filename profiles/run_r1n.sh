set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_n.log 2>&1; tail -5 gpurun_out/pytest_n.log
python bench.py --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; tail -c 800 gpurun_out/bench_n.err
