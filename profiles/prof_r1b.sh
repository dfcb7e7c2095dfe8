set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants"
$CMD > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'graph_sort_kernel|graph_edge_kernel|graph_resolve_kernel' -s 12 -c 3 -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out | tail -5
