set -x
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v.log 2>&1; tail -3 gpurun_out/pytest_u.log
timeout 60 python tools/sanitize_smoke.py > gpurun_out/smoke_v.log 2>&1; tail -1 gpurun_out/smoke_v.log
timeout 120 python bench.py --no-cpu-baseline --no-torch-gpu-baseline --no-other-configs > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo rc=$?; tail -c 300 gpurun_out/bench_v.err
