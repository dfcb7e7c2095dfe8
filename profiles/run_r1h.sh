set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_h.log 2>&1; tail -5 gpurun_out/pytest_h.log
python bench.py --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; tail -c 600 gpurun_out/bench_h.err
