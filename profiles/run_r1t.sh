# final state of round 1: GPU test suite + smoke + bench (own arm, reference arm)
set -x
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/pytest_t.log 2>&1; tail -2 gpurun_out/pytest_t.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_t.log 2>&1; tail -1 gpurun_out/smoke_t.log
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_t_ref.json 2> gpurun_out/bench_t.err
timeout 240 python bench.py > gpurun_out/bench_t.json 2>> gpurun_out/bench_t.err; echo rc=$?; tail -c 300 gpurun_out/bench_t.err
