set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-other-configs --no-torch-gpu-baseline"
$CMD > gpurun_out/plain_e.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'graph_edge_kernel|graph_sort_kernel|graph_gather_kernel' -s 15 -c 3 -o gpurun_out/prof_r1e $CMD > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out | tail -3
