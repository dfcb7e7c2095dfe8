set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_g.log 2>&1; tail -5 gpurun_out/pytest_g.log
python bench.py --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; tail -c 300 gpurun_out/bench_g.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-other-configs --no-torch-gpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'graph_edge_kernel|graph_sort_kernel' -s 10 -c 2 -o gpurun_out/prof_r1g $CMD > gpurun_out/ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'filter_emit_kernel|filter_count_kernel' -s 10 -c 2 -o gpurun_out/prof_r1g2 $CMD --nc 80 --conf 0.001 > gpurun_out/ncu_g2.log 2>&1
