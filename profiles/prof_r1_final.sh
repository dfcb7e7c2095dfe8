# round 1 final evidence: tests, bench (own arm + reference arm), ncu launch list, ncu --set full captures
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; tail -3 gpurun_out/pytest_final.log
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench.err
timeout 300 python bench.py > gpurun_out/final_bench.json 2>> gpurun_out/final_bench.err; tail -c 400 gpurun_out/final_bench.err
SHORT="timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-other-configs --no-torch-gpu-baseline"
$SHORT > gpurun_out/final_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv $SHORT > gpurun_out/final_ncu1.log 2>&1
K1='graph_sort_kernel|graph_gather_kernel|graph_edge_kernel|graph_resolve_kernel|loss_main_kernel|loss_positive_kernel|loss_finalize_kernel|filter_count_kernel|filter_emit_kernel'
ncu --set full --clock-control none --import-source on -k regex:"$K1" -s 54 -c 9 -o gpurun_out/final_prof $SHORT > gpurun_out/final_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'loss_main_kernel|filter_count_kernel|filter_emit_kernel' -s 18 -c 3 -o gpurun_out/final_prof_nc80 $SHORT --nc 80 --conf 0.001 > gpurun_out/final_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'loss_main_nchw_kernel|filter_count_nchw_kernel|filter_emit_nchw_kernel|assign_sparse_kernel' -s 24 -c 4 -o gpurun_out/final_prof_nc80_nchw $SHORT --nc 80 --conf 0.001 --layout nchw --targets labels > gpurun_out/final_ncu4.log 2>&1
ls -la gpurun_out | tail -12
