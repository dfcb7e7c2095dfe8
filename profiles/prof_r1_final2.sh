# refresh of the NMS evidence with the last edge kernel: launch list + full capture of the four graph kernels
set -x
SHORT="timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-other-configs --no-torch-gpu-baseline"
$SHORT > gpurun_out/final2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final2_launches.csv $SHORT > gpurun_out/final2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'graph_sort_kernel|graph_gather_kernel|graph_edge_kernel|graph_resolve_kernel' -s 24 -c 4 -o gpurun_out/final2_prof $SHORT > gpurun_out/final2_ncu2.log 2>&1
ls -la gpurun_out | tail -4
