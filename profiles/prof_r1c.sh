# round 1, session 2: tests + bench (all configs) + launch list + full capture of the HBM-bound kernels at nc=80
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_c.log 2>&1; tail -3 gpurun_out/pytest_c.log
python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -c 600 gpurun_out/bench_c.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c_ref.json 2>> gpurun_out/bench_c.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-other-configs --no-torch-gpu-baseline"
$CMD > gpurun_out/plain_c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c.csv $CMD > gpurun_out/ncu_c1.log 2>&1
CMD2="$CMD --nc 80 --conf 0.001"
$CMD2 > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'loss_main_kernel|loss_positive_kernel|filter_count_kernel|filter_emit_kernel' -s 12 -c 4 -o gpurun_out/prof_r1c $CMD2 > gpurun_out/ncu_c2.log 2>&1
ls -la gpurun_out | tail -12
