#!/bin/bash
# re-entry sanity pass after the build container was re-created: smoke(), the parity file without the full-size / fuzz cases, a short bench line
mkdir -p gpurun_out
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2w_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/r2w_smoke.log
( time timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not full_size and not fuzz and not config3 and not config2" ) > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2w_pytest.log
( time timeout 150 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-variants --no-torch-gpu-baseline ) > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
echo "bench rc=$?"; head -c 900 gpurun_out/r2w_bench.json
