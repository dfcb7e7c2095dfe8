set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f.log 2>&1; tail -5 gpurun_out/pytest_f.log
python bench.py --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -c 300 gpurun_out/bench_f.err
